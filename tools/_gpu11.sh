set -x
python bench.py --variant v2 --envs 4194304 --no-cpu-baseline > gpurun_out/r2_bench_v2_4m.json 2>gpurun_out/e1.log
python bench.py --variant v4 --envs 2097152 --no-cpu-baseline > gpurun_out/r2_bench_v4_2m.json 2>gpurun_out/e2.log
python bench.py --variant v5 --envs 1048576 --no-cpu-baseline > gpurun_out/r2_bench_v5_1m.json 2>gpurun_out/e3.log
python bench.py --variant v3 --envs 8388608 --window 1048576 --no-cpu-baseline > gpurun_out/r2_bench_v3_8m_window.json 2>gpurun_out/e4.log
python bench.py --variant v2 --envs 16777216 --obs-mode compact --no-cpu-baseline > gpurun_out/r2_bench_v2_compact.json 2>gpurun_out/e5.log
python bench.py --variant v4 --envs 8388608 --obs-mode compact --no-cpu-baseline > gpurun_out/r2_bench_v4_compact.json 2>gpurun_out/e6.log
python bench.py --variant v5 --envs 8388608 --obs-mode compact --no-cpu-baseline > gpurun_out/r2_bench_v5_compact.json 2>gpurun_out/e7.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_reference_arm.json 2>gpurun_out/e8.log
python bench.py > gpurun_out/r2_bench_n1.json 2>gpurun_out/e9.log
for f in gpurun_out/r2_bench_v*.json gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_n1_reference_arm.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d.get('roofline',{}); print('$f', round(d['value']/1e6,1),'M', round(d['e2e']['value']/1e6,1), r.get('achieved'), r.get('frac'), r.get('traffic'), d.get('clocks'), d.get('checks',{}).get('shard_invariance',{}).get('ok'))"; done
tail -2 gpurun_out/e*.log | grep -v "^$" | tail -20
