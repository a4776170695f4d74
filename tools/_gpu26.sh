python bench.py --variant v2 --envs 16777216 --obs-mode compact --steps 300 --no-cpu-baseline --no-extras > gpurun_out/x_v2c.json 2>gpurun_out/e8.log
python bench.py --variant v4 --envs 8388608 --obs-mode compact --steps 300 --no-cpu-baseline --no-extras > gpurun_out/x_v4c.json 2>gpurun_out/e5.log
for f in gpurun_out/x_v2c.json gpurun_out/x_v4c.json; do python -c "
import json,sys; d=json.load(open('$f')); r=d.get('roofline',{}); print('$f', round(d['value']/1e6,1),'M', round(d['e2e']['value']/1e6,1), d['warmup'], d['steps'], r['kernel_ms_min'], r['kernel_ms_median'], d.get('checks',{}).get('shard_invariance',{}).get('ok'))"; done
tail -n 3 gpurun_out/e8.log gpurun_out/e5.log
