for lib in default oldpre asmatom; do
  if [ $lib = default ]; then unset LMAZE_B200_LIB; else export LMAZE_B200_LIB=$PWD/build/liblmaze_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/r2_ab15.txt
  python tools/fov_sweep2.py v2 128x1 2>&1 | tee -a gpurun_out/r2_ab15.txt
  WARM=60 python tools/fov_sweep2.py v5,v4 128x1 2>&1 | tee -a gpurun_out/r2_ab15.txt
  WARM=60 python tools/fov_compact_sweep.py v4,v5 0 2>&1 | tee -a gpurun_out/r2_ab15.txt
  python tools/fov_compact_sweep.py v2 0 2>&1 | tee -a gpurun_out/r2_ab15.txt
  python bench.py --no-extras --no-cpu-baseline --steps 12 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('v0 headline', d['value']/1e6, 'M', d['roofline']['achieved'], 'GB/s e2e', d['e2e']['value']/1e6)" | tee -a gpurun_out/r2_ab15.txt
done
python tools/fov_sweep2.py v2 128x1 2>&1 | tee -a gpurun_out/r2_ab15.txt
