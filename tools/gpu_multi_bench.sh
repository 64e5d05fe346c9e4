# bench.py at N GPUs only (no tests).  usage: bash tools/gpu_multi_bench.sh TAG NGPU
TAG=${1:-mb}; N=${2:-4}; D=gpurun_out/$TAG; mkdir -p $D
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > $D/bench_n$N.json 2> $D/bench_n$N.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("$D/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms/step %.4f value %.4g" % (d["ms_per_step"], d["value"]), d.get("phases_ms_per_step"), d.get("verify"))
except Exception as e:
    print("bench parse failed", e); print(open("$D/bench_n$N.err").read()[-2500:])
PY
