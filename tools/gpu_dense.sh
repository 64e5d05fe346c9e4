TAG=${1:-dense}; N=${2:-4}
D=gpurun_out/$TAG; mkdir -p $D
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for kind in pore temp; do
timeout 400 $TR --master-port 29517 tests/nccl_slab_worker.py $D/dense_${kind}_n$N.json --kind $kind --mode p2p --dense --steps 30 > $D/dense_${kind}_n$N.log 2>&1; echo "exit $?" >> $D/dense_${kind}_n$N.log
done
timeout 600 $TR --master-port 29519 bench.py --gpus $N --steps 300 --warmup 5 > $D/bench300_n$N.json 2> $D/bench300_n$N.err; echo "exit $?" >> $D/bench300_n$N.err
