#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m multigpu -q --timeout=250 > gpurun_out/r2c27_pytest_multigpu.log 2>&1; echo "multigpu rc=$?"; tail -2 gpurun_out/r2c27_pytest_multigpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c27_bench_n2.json 2> gpurun_out/r2c27_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c27_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches')}, d['clocks']['reasons'])
r=d['ring']; print({k:v for k,v in r.items() if k not in ('variants','workload')}, list(r['variants']))
PY
