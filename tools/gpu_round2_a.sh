set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a/gpu.txt; nproc >> gpurun_out/r2a/gpu.txt
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a/gputests.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a/gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a/bench_default.json 2> gpurun_out/r2a/bench_default.err
python bench.py --workload cube --steps 20 --warmup 3 > gpurun_out/r2a/bench_cube_n1.json 2> gpurun_out/r2a/bench_cube_n1.err
python tools/bench_configs.py --steps 30 > gpurun_out/r2a/bench_configs.jsonl 2> gpurun_out/r2a/bench_configs.err
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py pore temp slab cube > gpurun_out/r2a/sanitizer_$tool.log 2>&1; echo "exit $?" >> gpurun_out/r2a/sanitizer_$tool.log
done
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a/bench_reference.json 2> gpurun_out/r2a/bench_reference.err
