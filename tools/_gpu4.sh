python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_hier.py -x -q -k "v4 or v5 or hier or visit or v2 or rollout or foveal" 2>&1 | tail -15 > gpurun_out/r2_t4.log; tail -7 gpurun_out/r2_t4.log
python tools/fov_sweep2.py v2,v4,v5 128x1,160x1,192x1,256x1 > gpurun_out/r2_fov_sweep3.txt 2>&1; cat gpurun_out/r2_fov_sweep3.txt
python tools/fov_compact_sweep.py v2,v4,v5 0,1,2 > gpurun_out/r2_fovc_sweep3.txt 2>&1; cat gpurun_out/r2_fovc_sweep3.txt
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 2 -c 1 -o gpurun_out/r2_ncu_v2_compact python tools/profile_one.py v2 compact 23 > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 2 -c 1 -o gpurun_out/r2_ncu_v4_compact2 python tools/profile_one.py v4 compact 22 > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmz_env_incr -s 1 -c 1 -o gpurun_out/r2_ncu_v0_incr python tools/profile_one.py v0 full 20 incremental 4 > gpurun_out/ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmz_env_fov -s 2 -c 1 -o gpurun_out/r2_ncu_v4_full2 python tools/profile_one.py v4 full 19 > gpurun_out/ncu6.log 2>&1
tail -n 2 gpurun_out/ncu3.log gpurun_out/ncu4.log gpurun_out/ncu5.log gpurun_out/ncu6.log
