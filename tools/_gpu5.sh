python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t5.log; tail -8 gpurun_out/r2_t5.log
python tools/fov_sweep2.py v4,v5 128x1,160x1,192x1,256x1 > gpurun_out/r2_fov_sweep4.txt 2>&1; cat gpurun_out/r2_fov_sweep4.txt
ncu --set full --clock-control none --import-source on -k regex:lmz_env_fov -s 2 -c 1 -o gpurun_out/r2_ncu_v4_full3 python tools/profile_one.py v4 full 19 > gpurun_out/ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmz_env_incr -s 1 -c 1 -o gpurun_out/r2_ncu_v0_incr python tools/profile_one.py v0 full 20 incremental 4 > gpurun_out/ncu5.log 2>&1
tail -n 2 gpurun_out/ncu5.log gpurun_out/ncu6.log
