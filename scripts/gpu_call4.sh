#!/bin/bash
# round 2, GPU call 4: bring-up of the fused D=128 backward (parity vs the two-kernel path, then timing), reference-error test
mkdir -p gpurun_out
timeout 600 python scripts/dev_fused.py check128 > gpurun_out/r2c4_check128.jsonl 2> gpurun_out/r2c4_check128.err; echo "check rc=$?"
cat gpurun_out/r2c4_check128.jsonl
if ! grep -q '"nan": true\|"rc"\|hang": \[' gpurun_out/r2c4_check128.jsonl; then
  timeout 300 python scripts/dev_fused.py bench128 > gpurun_out/r2c4_bench128.jsonl 2> gpurun_out/r2c4_bench128.err; echo "bench rc=$?"
  cat gpurun_out/r2c4_bench128.jsonl
fi
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "error_no_worse" > gpurun_out/r2c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c4_pytest.log
