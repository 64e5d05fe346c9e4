TAG=${1:-p}; N=${2:-4}
D=gpurun_out/$TAG; mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/slab_probe.py > $D/probe_n$N.log 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tests/nccl_slab_worker.py $D/worker_p2p_n$N.json --particles 16000000 --steps 12 --mode p2p > $D/worker_p2p_n$N.log 2>&1; echo "exit $?" >> $D/worker_p2p_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 3 > $D/bench_p2p_n$N.json 2> $D/bench_p2p_n$N.err; echo "exit $?" >> $D/bench_p2p_n$N.err
