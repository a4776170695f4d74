python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_hier.py -x -q -k "v4 or v5 or v6 or hier or visit or v2" 2>&1 | tail -8 > gpurun_out/r2_t17.log; tail -4 gpurun_out/r2_t17.log
WARM=60 python tools/fov_compact_sweep.py v4,v5 0,1,3 2>&1 | tee gpurun_out/r2_fovc_sweep_byte_steady.txt
python tools/fov_compact_sweep.py v2,v4,v5 0 2>&1 | tee -a gpurun_out/r2_fovc_sweep_byte_steady.txt
WARM=60 python tools/fov_sweep2.py v4,v5 128x1 2>&1 | tee gpurun_out/r2_fov_sweep_byte_steady.txt
python tools/fov_sweep2.py v2 128x1 2>&1 | tee -a gpurun_out/r2_fov_sweep_byte_steady.txt
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 70 -c 1 -o /tmp/v4c python tools/profile_one.py v4 compact 21 tma 72 > gpurun_out/ncu17.log 2>&1
ncu -i /tmp/v4c.ncu-rep --page details > gpurun_out/r2_ncu_v4c_steady4_details.txt 2>/dev/null
ncu -i /tmp/v4c.ncu-rep --page source --csv > gpurun_out/r2_ncu_v4c_steady4_source.csv 2>/dev/null
python tools/ncu_top_stalls.py gpurun_out/r2_ncu_v4c_steady4_source.csv 30 > gpurun_out/r2_ncu_v4c_steady4_stalls.txt; head -3 gpurun_out/r2_ncu_v4c_steady4_stalls.txt
