# event-driven cube sweep: parity + timing of both sweep variants.  usage: bash tools/gpu_sweep.sh TAG
TAG=${1:-sweep}; D=gpurun_out/$TAG; mkdir -p $D
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_operators.py tests/test_gpu_drivers.py tests/test_gpu_statistics.py -k "cube or operator or pairwise" -m gpu -q -x --durations=8 > $D/tests.log 2>&1; echo "pytest exit $?" >> $D/tests.log
timeout 300 python tools/bench_configs.py cfg1 --steps 50 > $D/cfg1_events.jsonl 2> $D/cfg1_events.err
AMC_CUBE_SWEEP=serial timeout 300 python tools/bench_configs.py cfg1 --steps 50 > $D/cfg1_serial.jsonl 2> $D/cfg1_serial.err
tail -5 $D/tests.log; cat $D/cfg1_events.jsonl $D/cfg1_serial.jsonl
