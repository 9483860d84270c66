#!/bin/bash
# round 2, GPU call 2: headline bench (both arms), TMEM shape probe, forward A/B, bf16 goldens, GPU test suite (bounded), ncu on C4
mkdir -p gpurun_out && rm -f gpurun_out/parity_errors.jsonl
timeout 300 python bench.py > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference > gpurun_out/r2c2_bench_ref.json 2> gpurun_out/r2c2_bench_ref.err; echo "bench ref rc=$?"
timeout 60 build/fa_probe_tmem > gpurun_out/r2c2_probe_tmem.txt 2>&1; echo "probe rc=$?"
for v in fwd_base fwd_max3 fwd_ld fwd_p default fwd_all_poly2 fwd_all_poly3; do
  if [ "$v" = default ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi
  timeout 120 python scripts/ab_time.py fwd >> gpurun_out/r2c2_ab_fwd.jsonl 2>> gpurun_out/r2c2_ab_fwd.err
done
unset FA_SM100_LIB
timeout 120 python tests/golden/make_golden_gpu.py > gpurun_out/r2c2_golden.log 2>&1; echo "golden rc=$?"
timeout 800 python -m pytest tests -m gpu -q -rfEs --timeout=200 --durations=15 > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c2_pytest.log
FA_BENCH_PREWARM_S=0 timeout 120 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c2_plain.log 2>&1 && \
FA_BENCH_PREWARM_S=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/r2c2_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c2_ncu1.log 2>&1
FA_BENCH_PREWARM_S=0 timeout 120 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c2_plain2.log 2>&1 && \
FA_BENCH_PREWARM_S=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -o gpurun_out/r2c2_prof_c4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c2_ncu2.log 2>&1
echo "ncu done rc=$?"
