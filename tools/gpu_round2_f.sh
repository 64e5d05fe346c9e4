set -x
D=gpurun_out/${1:-r2f}; mkdir -p $D
python tools/phase_clocks.py temp_scaled > $D/phase_temp.log 2>&1
python tools/phase_clocks.py pore_ref > $D/phase_pore.log 2>&1
python tools/profile_target.py pore_ref 6 > $D/plain_pore.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $D/launches_pore.csv python tools/profile_target.py pore_ref 6 > $D/ncu_pore.log 2>&1
