# usage: bash tools/gpu_variants.sh TAG name1 name2 ...   -- profile_target temp_scaled with libamc_<name>.so each ("base" = libamc.so)
TAG=$1; shift
D=gpurun_out/$TAG; mkdir -p $D
for v in "$@"; do
  if [ "$v" = base ]; then L=$PWD/argon_monte_carlo_b200/libamc.so; else L=$PWD/argon_monte_carlo_b200/libamc_$v.so; fi
  AMC_LIBRARY=$L python tools/profile_target.py temp_scaled 10 > $D/$v.log 2>&1
done
