python bench.py > gpurun_out/r2_bench_n1.json 2>gpurun_out/e1.log
python bench.py --variant v4 --envs 8388608 --obs-mode compact --steps 600 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v4_compact.json 2>gpurun_out/e5.log
python bench.py --variant v5 --envs 8388608 --obs-mode compact --steps 500 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v5_compact.json 2>gpurun_out/e6.log
python bench.py --variant v2 --envs 16777216 --obs-mode compact --steps 700 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_v2_compact.json 2>gpurun_out/e8.log
for f in gpurun_out/r2_bench_v*compact.json gpurun_out/r2_bench_n1.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d.get('roofline',{}); print('$f', round(d['value']/1e6,1),'M', round(d['e2e']['value']/1e6,1), r.get('achieved'), r.get('frac'), r.get('traffic'), d.get('clocks'), d.get('checks',{}).get('shard_invariance',{}).get('ok'))"; done
