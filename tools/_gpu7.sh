python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t7.log; tail -6 gpurun_out/r2_t7.log
python tools/small_bench.py 0 > gpurun_out/r2_small_bench2.txt 2>&1; cat gpurun_out/r2_small_bench2.txt
python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -3 gpurun_out/r2_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench2.json')); print(d['value'], d['e2e']['value'], d['checks']); print({k:(v.get('value') if isinstance(v,dict) else v) for k,v in d['extras'].items()})"
