set -x
D=gpurun_out/r2d; mkdir -p $D
for m in 0 1 2; do AMC_EXT_MODE=$m python tools/profile_target.py temp_scaled 8 > $D/mode$m.log 2>&1; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_phases_and_edges.py tests/test_gpu_init.py -m gpu -x -q > $D/gputests.log 2>&1; echo "pytest exit $?" >> $D/gputests.log
