# generic A/B: bench.py under different environments.  usage: bash tools/gpu_env_ab.sh TAG [tests] -- "name VAR=val ..." "name2 ..."
TAG=$1; shift; D=gpurun_out/$TAG; mkdir -p $D
if [ "$1" = "tests" ]; then
  shift
  timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
  tail -4 $D/gputests.log
fi
[ "$1" = "--" ] && shift
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-ref --no-verify --no-also > $D/bench_$name.json 2> $D/bench_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$D/bench_$name.json").read().strip().splitlines()[-1])
    print("$name ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["phases_ms_per_step"].items()}, "scatter %.4f" % d["roofline_longest_kernel"]["avg_launch_ms"], d["state_digest"]["digest"])
except Exception as e:
    print("$name bench parse failed", e); print(open("$D/bench_$name.err").read()[-1500:])
PY
done
