#!/bin/bash
# round 2, GPU call 8 (2 GPUs): multi-GPU tests, bench --gpus 2 (C4 strong-scaled + C5 ring block), new fused128 test
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c8_topo.txt 2>&1
timeout 100 python scripts/repro_map.py > gpurun_out/r2c8_repro.log 2>&1; head -3 gpurun_out/r2c8_repro.log
timeout 600 python -m pytest tests -m multigpu -q -rfEs --timeout=400 > gpurun_out/r2c8_pytest_multigpu.log 2>&1; echo "multigpu pytest rc=$?"; tail -5 gpurun_out/r2c8_pytest_multigpu.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "fused128" > gpurun_out/r2c8_pytest_fused128.log 2>&1; echo "fused128 pytest rc=$?"; tail -3 gpurun_out/r2c8_pytest_fused128.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c8_bench_n2.json 2> gpurun_out/r2c8_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2c8_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c8_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'] and {k:v for k,v in d['e2e'].items() if k!='note'})
r=d.get('ring'); 
if r:
    print({k:v for k,v in r.items() if k not in('hops_rank0','note')})
    for h in r.get('hops_rank0',[]): print(h)
for k,v in (d.get('also') or {}).items(): print(k, v['ms_per_step'], v['fwd_bwd_tflops'], v['host_us_per_step'])
PY
