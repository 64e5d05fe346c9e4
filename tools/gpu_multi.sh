# usage: bash tools/gpu_multi.sh TAG NGPU [steps] [particles for the worker]
set -x
TAG=${1:-m}; N=${2:-2}; K=${3:-20}; M=${4:-4000000}
D=gpurun_out/$TAG; mkdir -p $D
python -m pytest tests/test_gpu_slab.py tests/test_gpu_nccl_slab.py -m gpu -x -q > $D/tests.log 2>&1; echo "pytest exit $?" >> $D/tests.log
python tools/profile_target.py slab1 8 > $D/slab1.log 2>&1
for mode in ${MODES:-p2p}; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/nccl_slab_worker.py $D/worker_${mode}_n$N.json --particles $M --steps 12 --mode $mode > $D/worker_${mode}_n$N.log 2>&1; echo "exit $?" >> $D/worker_${mode}_n$N.log
AMC_SLAB_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps $K --warmup 3 > $D/bench_${mode}_n$N.json 2> $D/bench_${mode}_n$N.err; echo "exit $?" >> $D/bench_${mode}_n$N.err
done
