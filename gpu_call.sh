mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c32_pytest.log
timeout 200 python scripts/ab_time.py all 2>/dev/null | cut -c1-400
timeout 300 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {k:round(v['ms'],4) for k,v in d['kernels'].items()})"
