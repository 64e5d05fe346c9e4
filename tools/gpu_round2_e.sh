set -x
D=gpurun_out/${1:-r2e}; mkdir -p $D
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_phases_and_edges.py -m gpu -x -q > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
python tools/profile_target.py temp_scaled 8 > $D/plain8.log 2>&1
python tools/profile_target.py temp_scaled 4 > $D/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_detect" -s 2 -c 1 -f -o $D/prof python tools/profile_target.py temp_scaled 4 > $D/ncu.log 2>&1
echo "ncu exit $?" >> $D/ncu.log
