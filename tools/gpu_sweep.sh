D=gpurun_out/${1:-sw}; mkdir -p $D
for v in base sw256 sw128 sw64; do
  if [ "$v" = base ]; then L=$PWD/argon_monte_carlo_b200/libamc.so; else L=$PWD/argon_monte_carlo_b200/libamc_$v.so; fi
  AMC_LIBRARY=$L python tools/bench_configs.py cfg1 --steps 30 > $D/cfg1_$v.json 2> $D/cfg1_$v.err
done
