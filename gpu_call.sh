timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/c46_bench.json 2> gpurun_out/c46_bench.err; tail -2 gpurun_out/c46_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c46_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print({k:(round(v['ms'],4), round(v['achieved'],1)) for k,v in d['kernels'].items()})
print(d['roofline'])
print({k:(round(v['fwd_tflops']), round(v['fwd_bwd_tflops'])) for k,v in d['also'].items()})
PY
timeout 300 python bench.py --workload C3 > gpurun_out/c46_bench_c3.json 2>> gpurun_out/c46_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c46_bench_c3.json')); print('C3', d['value'], d['ms_per_step'], {k:(round(v['ms'],4), round(v['achieved'],1)) for k,v in d['kernels'].items()})"
