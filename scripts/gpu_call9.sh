#!/bin/bash
# round 2, GPU call 9 (1 GPU): gather variant with the real kernels (virtual ranks), per-rank C5 compute (gather vs ring), the
# like-for-like sweep vs the reference, ncu launch list of the headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -k "gather or fused128" --timeout=300 > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c9_pytest.log
for r in 3 0 7; do timeout 300 python scripts/bench_cp_local.py $r 4 >> gpurun_out/r2c9_cp_local.jsonl 2>> gpurun_out/r2c9_cp_local.err; done
cat gpurun_out/r2c9_cp_local.jsonl; tail -3 gpurun_out/r2c9_cp_local.err
timeout 900 python scripts/sweep_vs_reference.py --out gpurun_out/r02_sweep_vs_reference.jsonl > gpurun_out/r02_sweep_vs_reference.md 2> gpurun_out/r2c9_sweep.err; echo "sweep rc=$?"
cat gpurun_out/r02_sweep_vs_reference.md
FA_BENCH_PREWARM_S=0 timeout 120 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c9_plain.log 2>&1 && \
FA_BENCH_PREWARM_S=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_ -c 40 --csv --log-file gpurun_out/r2c9_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c9_ncu1.log 2>&1
echo "ncu rc=$?"
