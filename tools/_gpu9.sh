python -m pytest tests/test_integration_stub.py -x -q 2>&1 | tail -3
ncu --set full --clock-control none -k regex:lmz_ -o gpurun_out/r2_final python tools/profile_final.py gpurun_out/r2_final_manifest.json > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i gpurun_out/r2_final.ncu-rep --page raw --csv > gpurun_out/r2_ncu_final_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_final.ncu-rep --page details > gpurun_out/r2_ncu_final_details.txt 2>/dev/null
rm -f gpurun_out/r2_final.ncu-rep
ls -la gpurun_out/r2_ncu_final_raw.csv gpurun_out/r2_ncu_final_details.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b3.json 2>/dev/null && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v0_tma_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-300
