python -m pytest tests/test_gpu_scale.py tests/test_gpu_r2.py tests/test_gpu_parity.py tests/test_gpu_hier.py -x -q -k "scale or million or compact or soak or foveal or hier_batched or windows" 2>&1 | tail -15 > gpurun_out/r2_t6.log; tail -8 gpurun_out/r2_t6.log
python tools/fov_compact_sweep.py v2,v4,v5 0,1,2 > gpurun_out/r2_fovc_sweep4.txt 2>&1; cat gpurun_out/r2_fovc_sweep4.txt
for v in noprev nowb noprevwb novisit; do echo "== $v"; LMAZE_B200_LIB=$PWD/build/liblmaze_$v.so python tools/fov_sweep2.py v4,v5 128x1; done > gpurun_out/r2_fov_variants.txt 2>&1; cat gpurun_out/r2_fov_variants.txt
