#!/bin/bash
mkdir -p gpurun_out
echo skip
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c30_bench_n2.json 2> gpurun_out/r2c30_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c30_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling','gpu_launches')}, d['clocks']['reasons'])
r=d['ring']; print({k:v for k,v in r.items() if k not in ('variants','workload')}, list(r['variants']))
PY
