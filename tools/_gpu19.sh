python tools/fov_compact_sweep.py v2 0 2>&1 | tee gpurun_out/r2_v2c_check.txt
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 141 -c 1 -o /tmp/v5c python tools/profile_one.py v5 compact 21 tma 72 > gpurun_out/ncu19.log 2>&1
ncu -i /tmp/v5c.ncu-rep --page details > gpurun_out/r2_ncu_v5c_steady_details.txt 2>/dev/null
ncu -i /tmp/v5c.ncu-rep --page source --csv > gpurun_out/r2_ncu_v5c_steady_source.csv 2>/dev/null
python tools/ncu_top_stalls.py gpurun_out/r2_ncu_v5c_steady_source.csv 40 > gpurun_out/r2_ncu_v5c_steady_stalls.txt; head -3 gpurun_out/r2_ncu_v5c_steady_stalls.txt
grep -E "Duration|Issue Slots Busy|Grid Size|Registers Per" gpurun_out/r2_ncu_v5c_steady_details.txt
