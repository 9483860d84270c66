#!/bin/bash
# round 2, GPU call 11 (8 GPUs): bench --gpus 8 (C4 strong-scaled + C5 both variants), multi-GPU tests
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2c21_bench_n8.json 2> gpurun_out/r2c21_bench_n8.err; echo "bench n8 rc=$?"; grep "ring_block rank 0" gpurun_out/r2c21_bench_n8.err | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c21_bench_n8.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','scaling')}, d['e2e'] and {k:v for k,v in d['e2e'].items() if k!='note'})
r=d.get('ring')
if r:
    print({k:v for k,v in r.items() if k not in('variants','note')})
    for n,v in r.get('variants',{}).items():
        print(n, {k:x for k,x in v.items() if k not in('hops_rank0','segments_rank0','note')})
        for h in v.get('hops_rank0',[])+v.get('segments_rank0',[]): print('   ',h)
for k,v in (d.get('also') or {}).items(): print(k, v['ms_per_step'], v['fwd_bwd_tflops'], v['host_us_per_step'])
PY
echo skip-multigpu-tests
