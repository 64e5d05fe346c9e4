# quick validation: full GPU suite + a short bench without the CPU legs.  usage: bash tools/gpu_quick.sh TAG [pytest -k expr]
TAG=${1:-quick}; D=gpurun_out/$TAG; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q -x --durations=10 ${2:+-k "$2"} > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-ref --no-verify > $D/bench_quick.json 2> $D/bench_quick.err
tail -15 $D/gputests.log; python - <<PY
import json
try:
    d = json.loads(open("$D/bench_quick.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], d["phases_ms_per_step"], "detect frac", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], "scatter", d["roofline_longest_kernel"]["avg_launch_ms"], "also", d.get("also", {}).get("ms_per_step"), [c["ms_per_step"] for c in d.get("also_configs", [])])
except Exception as e:
    print("bench parse failed", e); print(open("$D/bench_quick.err").read()[-2000:])
PY
