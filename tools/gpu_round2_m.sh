set -x
D=gpurun_out/${1:-r2m}; mkdir -p $D
for occ in 4 6; do AMC_LIBRARY=$PWD/argon_monte_carlo_b200/libamc_occ$occ.so python tools/profile_target.py slab1 8 > $D/slab1_occ$occ.log 2>&1; done
python tools/profile_target.py slab1 4 > $D/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_keys" -s 2 -c 1 -f -o $D/prof python tools/profile_target.py slab1 4 > $D/ncu.log 2>&1
