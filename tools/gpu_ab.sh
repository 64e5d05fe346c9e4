# A/B of the detection pass.  usage: bash tools/gpu_ab.sh TAG "modes" [tests]   (tests: run the full GPU suite first)
TAG=${1:-ab}; MODES=${2:-"tma tma8 ldg"}; D=gpurun_out/$TAG; mkdir -p $D
if [ -n "$3" ]; then
  timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
  tail -4 $D/gputests.log
fi
for mode in $MODES; do
  AMC_DEBUG=1 AMC_DETECT=$mode timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-ref --no-verify --no-also > $D/bench_$mode.json 2> $D/bench_$mode.err
  grep "amc:" $D/bench_$mode.err | head -1
  python - <<PY
import json
try:
    d = json.loads(open("$D/bench_$mode.json").read().strip().splitlines()[-1])
    print("$mode ms/step %.4f" % d["ms_per_step"], {k: round(v, 4) for k, v in d["phases_ms_per_step"].items()}, "detect frac %.3f" % d["roofline"]["frac"], d["state_digest"]["digest"])
except Exception as e:
    print("$mode bench parse failed", e); print(open("$D/bench_$mode.err").read()[-1500:])
PY
done
