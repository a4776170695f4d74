python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_hier.py tests/test_gpu_compat.py -x -q -k "v4 or v5 or v6 or hier or visit or foveal or rollout or compat" 2>&1 | tail -15 > gpurun_out/r2_t12.log; tail -8 gpurun_out/r2_t12.log
python tools/fov_sweep2.py v4,v5 128x1,160x1,192x1,256x1 2>&1 | tee gpurun_out/r2_fov_sweep_hist.txt
python tools/fov_compact_sweep.py v4,v5 0,1,2,3 2>&1 | tee gpurun_out/r2_fovc_sweep_hist.txt
python tools/rollout_bench2.py 2>&1 | tee gpurun_out/r2_rollout_hist.txt
