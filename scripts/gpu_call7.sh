#!/bin/bash
# round 2, GPU call 7: speculative-max forward: parity (GPU test suite) + A/B timing
mkdir -p gpurun_out
for v in fwd_spec0 default fwd_spec2 fwd_spec1_poly2 fwd_spec1_poly0; do
  if [ "$v" = default ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi
  timeout 120 python scripts/ab_time.py fwd >> gpurun_out/r2c7_ab_fwd.jsonl 2>> gpurun_out/r2c7_ab_fwd.err
done
unset FA_SM100_LIB
cat gpurun_out/r2c7_ab_fwd.jsonl
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c7_pytest.log
