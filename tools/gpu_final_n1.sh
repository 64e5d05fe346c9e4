set -x
D=gpurun_out/${1:-fin1}; mkdir -p $D
python -m pytest tests -m gpu -q --durations=10 > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
python bench.py > $D/bench_default.json 2> $D/bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > $D/bench_reference.json 2> $D/bench_reference.err
python bench.py --workload cube --steps 20 --warmup 3 > $D/bench_cube_n1.json 2> $D/bench_cube_n1.err
python tools/bench_configs.py cfg1 cfg2 cfg3 cfg3host --steps 30 > $D/bench_configs.jsonl 2> $D/bench_configs.err
python bench.py --steps 2 --warmup 3 --no-also --no-cpu --no-ref --no-verify > $D/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu --no-ref --no-verify > $D/ncu_list.log 2>&1
python tools/profile_target.py temp_scaled 4 > $D/plain_pt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_detect|k_keys|k_scatter_advect|k_pairs_group|k_build_worklist" -s 30 -c 12 -f -o $D/prof python tools/profile_target.py temp_scaled 4 > $D/ncu_full.log 2>&1
echo "ncu exit $?" >> $D/ncu_full.log
