#!/bin/bash
# GPU call 13 (2 GPUs): dynamic first item (shared-SM mode) in the sequence-parallel paths
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q -k "shared_sm or virtual" --timeout=200 > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c13_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c13_bench_n2.json 2> gpurun_out/r2c13_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2c13_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c13_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})
r=d.get('ring')
if r:
    print({k:v for k,v in r.items() if k not in('variants','note')})
    for n,v in r.get('variants',{}).items():
        print(n, {k:x for k,x in v.items() if k not in('hops_rank0','segments_rank0','note')})
        for h in v.get('segments_rank0',[])+v.get('hops_rank0',[]): print('   ',h)
PY
