for lib in default v2regs; do
  if [ $lib = default ]; then unset LMAZE_B200_LIB; else export LMAZE_B200_LIB=$PWD/build/liblmaze_$lib.so; fi
  echo "== $lib" | tee -a gpurun_out/r2_ab18.txt
  python tools/fov_compact_sweep.py v2 0,1,2,3 2>&1 | tee -a gpurun_out/r2_ab18.txt
  python tools/fov_compact_sweep.py v2 1,2 64 2>&1 | tee -a gpurun_out/r2_ab18.txt
  python tools/fov_compact_sweep.py v2 1,2 256 2>&1 | tee -a gpurun_out/r2_ab18.txt
done
unset LMAZE_B200_LIB
python -m pytest tests/test_gpu_r2.py tests/test_gpu_parity.py -x -q -k "v2 or foveal" 2>&1 | tail -4
