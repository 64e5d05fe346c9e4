# usage: bash tools/gpu_ncu.sh TAG [kernel regex]   -- plain run first, then one ncu --set full capture of the matching kernels of step 3
set -x
TAG=${1:-x}; RE=${2:-"k_detect|k_keys|k_scatter_advect"}
D=gpurun_out/$TAG; mkdir -p $D
python tools/profile_target.py temp_scaled 4 > $D/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 6 -c 3 -f -o $D/prof python tools/profile_target.py temp_scaled 4 > $D/ncu.log 2>&1
echo "ncu exit $?" >> $D/ncu.log
