#!/bin/bash
# GPU call 12 (2 GPUs): gather variant with uneven head groups / capped NCCL group: functional check + timing at P=2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c12_bench_n2.json 2> gpurun_out/r2c12_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2c12_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c12_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})
r=d.get('ring')
if r:
    print({k:v for k,v in r.items() if k not in('variants','note')})
    for n,v in r.get('variants',{}).items():
        print(n, {k:x for k,x in v.items() if k not in('hops_rank0','segments_rank0','note')})
        for h in v.get('segments_rank0',[]): print('   ',h)
PY
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "without_cuda_context" > gpurun_out/r2c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c12_pytest.log
timeout 300 python -m pytest tests -m multigpu -q --timeout=300 > gpurun_out/r2c12_pytest_multigpu.log 2>&1; echo "multigpu rc=$?"; tail -3 gpurun_out/r2c12_pytest_multigpu.log
