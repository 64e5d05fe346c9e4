set -x
D=gpurun_out/${1:-r2j}; mkdir -p $D
python tools/profile_target.py slab1 6 > $D/plain_slab1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file $D/launches_slab1.csv python tools/profile_target.py slab1 6 > $D/ncu_slab1.log 2>&1
python tools/profile_target.py temp_scaled 6 > $D/plain_temp.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 50 --csv --log-file $D/launches_temp.csv python tools/profile_target.py temp_scaled 6 > $D/ncu_temp.log 2>&1
