#!/bin/bash
# round 2, GPU call 5: where does the fused D=128 backward lose its time?  skip-experiments + ncu (full set, source page) on C3
mkdir -p gpurun_out
for v in default f128_skip1 f128_skip2 f128_skip15 f128_skip16 f128_skip31 f128_dkfirst f128_stagger f128_poly4; do
  if [ "$v" = default ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi
  timeout 120 python scripts/ab_fused128.py >> gpurun_out/r2c5_ab_fused128.jsonl 2>> gpurun_out/r2c5_ab_fused128.err
done
unset FA_SM100_LIB
cat gpurun_out/r2c5_ab_fused128.jsonl
timeout 120 python scripts/prof_fused.py 4 16 4096 0 128 > gpurun_out/r2c5_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused128 -s 2 -c 1 -o gpurun_out/r2c5_prof_fused128 python scripts/prof_fused.py 4 16 4096 0 128 > gpurun_out/r2c5_ncu.log 2>&1
echo "ncu rc=$?"
