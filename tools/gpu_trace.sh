set -x
TAG=${1:-t}; N=${2:-4}
D=gpurun_out/$TAG; mkdir -p $D
# one step per call so that every step is traced
AMC_SLAB_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 --no-verify > $D/bench_trace_n$N.json 2> $D/bench_trace_n$N.err
