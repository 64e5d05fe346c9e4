# usage: bash tools/gpu_multi8.sh TAG NGPU [steps]   -- worker parity check + bench at N GPUs (p2p mode)
set -x
TAG=${1:-m}; N=${2:-8}; K=${3:-20}
D=gpurun_out/$TAG; mkdir -p $D
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/nccl_slab_worker.py $D/worker_p2p_n$N.json --particles 16000000 --steps 12 --mode p2p > $D/worker_p2p_n$N.log 2>&1; echo "exit $?" >> $D/worker_p2p_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps $K --warmup 3 > $D/bench_p2p_n$N.json 2> $D/bench_p2p_n$N.err; echo "exit $?" >> $D/bench_p2p_n$N.err
