set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t22.log; tail -3 gpurun_out/r2_t22.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/r2_smoke22.log
ncu --set full --clock-control none --profile-from-start off -k regex:lmz_ -o /tmp/r2_final python tools/profile_final.py gpurun_out/r2_final_manifest.json > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i /tmp/r2_final.ncu-rep --page raw --csv > gpurun_out/r2_ncu_final_raw.csv 2>/dev/null
ncu -i /tmp/r2_final.ncu-rep --page details > gpurun_out/r2_ncu_final_details.txt 2>/dev/null
ls -la gpurun_out/r2_ncu_final_raw.csv
python bench.py > gpurun_out/r2_bench_n1.json 2>gpurun_out/e1.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_reference_arm.json 2>gpurun_out/e2.log
python bench.py --variant v4 --envs 2097152 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v4_2m.json 2>gpurun_out/e3.log
python bench.py --variant v5 --envs 1048576 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v5_1m.json 2>gpurun_out/e4.log
python bench.py --variant v4 --envs 8388608 --obs-mode compact --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v4_compact.json 2>gpurun_out/e5.log
python bench.py --variant v5 --envs 8388608 --obs-mode compact --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v5_compact.json 2>gpurun_out/e6.log
python bench.py --variant v2 --envs 4194304 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_4m.json 2>gpurun_out/e7.log
python bench.py --variant v2 --envs 16777216 --obs-mode compact --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_compact.json 2>gpurun_out/e8.log
python bench.py --variant v3 --envs 8388608 --window 1048576 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v3_8m_window.json 2>gpurun_out/e9.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_v0_tma_bench_steps3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
set +x
for f in gpurun_out/r2_bench_v*.json gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_n1_reference_arm.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d.get('roofline',{}); print('$f', round(d['value']/1e6,1),'M', round(d['e2e']['value']/1e6,1), r.get('achieved'), r.get('frac'), r.get('traffic'), (d.get('clocks') or {}).get('sm_mhz'), d.get('checks',{}).get('shard_invariance',{}).get('ok'))"; done
tail -2 gpurun_out/e*.log | grep -v "^$" | tail -20
