set -x
python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_hier.py -x -q -k "compact or incremental or hier or v4 or v2" 2>&1 | tail -8 > gpurun_out/r2_t3.log; tail -5 gpurun_out/r2_t3.log
python tools/fov_compact_sweep.py v2,v4,v5 0,1,2,3 > gpurun_out/r2_fovc_sweep2.txt 2>&1; cat gpurun_out/r2_fovc_sweep2.txt
python tools/small_bench.py 0 > gpurun_out/r2_small_bench.txt 2>&1; cat gpurun_out/r2_small_bench.txt
LMAZE_B200_LIB=$PWD/build/liblmaze_novisit.so python tools/fov_sweep2.py v4,v5 128x1,160x1,224x1,256x1 > gpurun_out/r2_fov_novisit.txt 2>&1; cat gpurun_out/r2_fov_novisit.txt
ncu --set full --clock-control none --import-source on -k regex:lmz_fov_small -s 2 -c 1 -o gpurun_out/r2_ncu_v4_compact python tools/profile_one.py v4 compact 22 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lmz_env_fov -s 2 -c 1 -o gpurun_out/r2_ncu_v4_full python tools/profile_one.py v4 full 19 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu1.log gpurun_out/ncu2.log
