# usage: bash tools/gpu_final_multi.sh TAG NGPU [ref]  -- the driver's command lines at N GPUs: pore bench, cube bench, NCCL-vs-single parity worker
set -x
TAG=${1:-finm}; N=${2:-2}
D=gpurun_out/$TAG; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 800 $TR --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > $D/bench_n$N.json 2> $D/bench_n$N.err; echo "exit $?" >> $D/bench_n$N.err
timeout 400 $TR --master-port 29517 tests/nccl_slab_worker.py $D/worker_p2p_n$N.json --particles 16000000 --steps 12 --mode p2p > $D/worker_p2p_n$N.log 2>&1; echo "exit $?" >> $D/worker_p2p_n$N.log
timeout 600 $TR --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --workload cube > $D/bench_cube_n$N.json 2> $D/bench_cube_n$N.err; echo "exit $?" >> $D/bench_cube_n$N.err
if [ "$3" = ref ]; then
timeout 800 $TR --master-port 29525 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $D/bench_reference_n$N.json 2> $D/bench_reference_n$N.err; echo "exit $?" >> $D/bench_reference_n$N.err
fi
