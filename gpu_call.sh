timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | cut -c1-300
timeout 300 python bench.py > gpurun_out/c58_bench.json 2> gpurun_out/c58_bench.err; tail -2 gpurun_out/c58_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c58_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks']['sm_mhz'])
print({k:(round(v['ms'],4), round(v['achieved'],1)) for k,v in d['kernels'].items()})
PY
