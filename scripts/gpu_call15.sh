#!/bin/bash
# GPU call 15 (2 GPUs): PeerCollectives (symmetric memory + copy engines): multi-GPU tests, then the bench ring block
mkdir -p gpurun_out
echo skip-tests
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2c20_bench_n2.json 2> gpurun_out/r2c20_bench_n2.err; echo "bench n2 rc=$?"; grep "ring_block rank 0" gpurun_out/r2c20_bench_n2.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c20_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')})
r=d.get('ring')
if r:
    print({k:v for k,v in r.items() if k not in('variants','note')})
    for n,v in r.get('variants',{}).items():
        print(n, {k:x for k,x in v.items() if k not in('hops_rank0','segments_rank0','note')})
        for h in v.get('segments_rank0',[]): print('   ',h)
PY
