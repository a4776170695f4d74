set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t27.log; tail -2 gpurun_out/r2_t27.log
grep -q " passed" gpurun_out/r2_t27.log && ! grep -q "failed" gpurun_out/r2_t27.log || exit 1
ncu --set full --clock-control none --profile-from-start off -k regex:lmz_ -o /tmp/r2_final python tools/profile_final.py gpurun_out/r2_final_manifest.json > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i /tmp/r2_final.ncu-rep --page raw --csv > gpurun_out/r2_ncu_final_raw.csv 2>/dev/null
ncu -i /tmp/r2_final.ncu-rep --page details > gpurun_out/r2_ncu_final_details.txt 2>/dev/null
python bench.py --variant v2 --envs 16777216 --obs-mode compact --steps 300 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_compact.json 2>gpurun_out/e8.log
python bench.py --variant v4 --envs 8388608 --obs-mode compact --steps 300 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v4_compact.json 2>gpurun_out/e5.log
python bench.py --variant v5 --envs 8388608 --obs-mode compact --steps 300 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v5_compact.json 2>gpurun_out/e6.log
python bench.py --variant v4 --envs 2097152 --steps 60 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v4_2m.json 2>gpurun_out/e3.log
python bench.py --variant v5 --envs 1048576 --steps 60 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v5_1m.json 2>gpurun_out/e4.log
python bench.py --variant v2 --envs 4194304 --steps 60 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_4m.json 2>gpurun_out/e7.log
set +x
for f in gpurun_out/r2_bench_v[245]*.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d.get('roofline',{}); print('$f', round(d['value']/1e6,1),'M', round(d['e2e']['value']/1e6,1), d['warmup'], d['steps'], round(r.get('achieved'),0), round(r.get('frac'),3), r['kernel_ms_min'], r['kernel_ms_median'], (d.get('clocks') or {}).get('samples'), d.get('checks',{}).get('shard_invariance',{}).get('ok'))"; done
