set -x
D=gpurun_out/r2b; mkdir -p $D
python -m pytest tests -m gpu -x -q --durations=10 > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
python bench.py --steps 20 --warmup 3 --no-ref --no-cpu > $D/bench.json 2> $D/bench.err
python tools/bench_configs.py cfg1 cfg2 cfg3 cfg3host --steps 30 > $D/bench_configs.jsonl 2> $D/bench_configs.err
