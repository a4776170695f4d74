python tools/sustained_probe.py v2 800 2>&1 | tee gpurun_out/r2_sustained_probe.txt
python tools/sustained_probe.py v0 800 2>&1 | tee -a gpurun_out/r2_sustained_probe.txt
python tools/sustained_probe.py v4 600 2>&1 | tee -a gpurun_out/r2_sustained_probe.txt
