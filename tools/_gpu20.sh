python -m pytest tests/test_gpu_hier.py tests/test_gpu_r2.py -x -q -k "v5 or v6 or hier" 2>&1 | tail -3
WARM=60 python tools/fov_compact_sweep.py v5,v4 0 2>&1 | tee gpurun_out/r2_fovc_sweep_runs_steady.txt
python tools/hier_bench.py 1048576 128 2>&1 | tail -3 | tee gpurun_out/r2_hier_bench_steady.txt
