python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -x -q -k "incremental" 2>&1 | tail -5 > gpurun_out/r2_t8.log; tail -3 gpurun_out/r2_t8.log
python tools/small_bench.py 0 2>&1 | grep incr
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], json.dumps(d['checks'])[:1500]); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d['extras'].items()}); print(json.dumps(d['extras']['rollout_T64']['checks']))"
