timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | cut -c1-250
timeout 100 python scripts/bench_varlen.py > gpurun_out/c66_varlen.jsonl 2>&1; cut -c1-330 gpurun_out/c66_varlen.jsonl
timeout 200 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('C2', round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches'])"
