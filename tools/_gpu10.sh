python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t10.log; tail -4 gpurun_out/r2_t10.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --set full --clock-control none -k regex:lmz_ -o gpurun_out/r2_final python tools/profile_final.py gpurun_out/r2_final_manifest.json > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i gpurun_out/r2_final.ncu-rep --page raw --csv > gpurun_out/r2_ncu_final_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_final.ncu-rep --page details > gpurun_out/r2_ncu_final_details.txt 2>/dev/null
rm -f gpurun_out/r2_final.ncu-rep
ls -la gpurun_out/r2_ncu_final_raw.csv
