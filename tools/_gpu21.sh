ncu --metrics gpu__time_duration.sum --clock-control none -k regex:lmz_ -s 139 -c 6 --csv --log-file gpurun_out/r2_v5c_launch_times.csv python tools/profile_one.py v5 compact 21 tma 73 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2_v5c_launch_times.csv | tail -7
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 141 -c 1 -o /tmp/v5c python tools/profile_one.py v5 compact 21 tma 72 > gpurun_out/ncu21.log 2>&1
ncu -i /tmp/v5c.ncu-rep --page details > gpurun_out/r2_ncu_v5c_planner2_details.txt 2>/dev/null
ncu -i /tmp/v5c.ncu-rep --page source --csv > gpurun_out/r2_ncu_v5c_planner2_source.csv 2>/dev/null
python tools/ncu_top_stalls.py gpurun_out/r2_ncu_v5c_planner2_source.csv 40 > gpurun_out/r2_ncu_v5c_planner2_stalls.txt; head -3 gpurun_out/r2_ncu_v5c_planner2_stalls.txt
grep -E "Duration|Issue Slots Busy" gpurun_out/r2_ncu_v5c_planner2_details.txt
