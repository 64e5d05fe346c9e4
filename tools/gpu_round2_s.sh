set -x
D=gpurun_out/${1:-r2s}; mkdir -p $D
python -m pytest tests/test_gpu_operators.py tests/test_gpu_init.py tests/test_gpu_parity.py tests/test_gpu_phases_and_edges.py tests/test_gpu_statistics.py tests/test_gpu_slab.py -m gpu -q --durations=8 > $D/tests.log 2>&1; echo "pytest exit $?" >> $D/tests.log
python tools/profile_target.py temp_scaled 8 > $D/temp.log 2>&1
python tools/profile_target.py pore_ref 8 > $D/pore.log 2>&1
