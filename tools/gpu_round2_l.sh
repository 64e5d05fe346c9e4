set -x
D=gpurun_out/${1:-r2l}; mkdir -p $D
python -m pytest tests/test_gpu_slab.py -m gpu -x -q > $D/tests.log 2>&1; echo "pytest exit $?" >> $D/tests.log
python tools/profile_target.py slab1 8 > $D/slab1.log 2>&1
python tools/profile_target.py temp_scaled 8 > $D/temp.log 2>&1
