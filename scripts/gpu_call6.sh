#!/bin/bash
# round 2, GPU call 6: fused D=128 backward with the dQ staging moved into the dead Q stage: parity, A/B, bench
mkdir -p gpurun_out
timeout 600 python scripts/dev_fused.py check128 > gpurun_out/r2c6_check128.jsonl 2> gpurun_out/r2c6_check128.err; echo "check rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2c6_check128.jsonl'):
    d=json.loads(l)
    if 'shape' in d: print(d['shape'], d['causal'], 'hang', d['hang'], {k:(d[k]['maxdiff'], d[k]['nan']) for k in ('dQ','dK','dV')})
    else: print(d)
PY
for v in default f128_skip1 f128_skip15 f128_stagger f128_poly4 f128_poly0; do
  if [ "$v" = default ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi
  timeout 120 python scripts/ab_fused128.py >> gpurun_out/r2c6_ab_fused128.jsonl 2>> gpurun_out/r2c6_ab_fused128.err
done
unset FA_SM100_LIB
cat gpurun_out/r2c6_ab_fused128.jsonl
timeout 300 python scripts/dev_fused.py bench128 > gpurun_out/r2c6_bench128.jsonl 2> gpurun_out/r2c6_bench128.err; echo "bench rc=$?"
cat gpurun_out/r2c6_bench128.jsonl
timeout 120 python scripts/prof_fused.py 4 16 4096 0 128 > gpurun_out/r2c6_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused128 -s 2 -c 1 -o gpurun_out/r2c6_prof_fused128 python scripts/prof_fused.py 4 16 4096 0 128 > gpurun_out/r2c6_ncu.log 2>&1
echo "ncu rc=$?"
