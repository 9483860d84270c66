#!/bin/bash
# round 2, GPU call 3 (after container re-creation): smoke, GPU test suite, headline bench (both arms), forward A/B, bf16 goldens, ncu on C4
mkdir -p gpurun_out && rm -f gpurun_out/parity_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c3_env.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c3_smoke.log
timeout 900 python -m pytest tests -m gpu -q -rfEs --timeout=240 --durations=15 > gpurun_out/r2c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2c3_pytest.log
timeout 400 python bench.py > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c3_bench.err
timeout 500 python bench.py --impl reference > gpurun_out/r2c3_bench_ref.json 2> gpurun_out/r2c3_bench_ref.err; echo "bench ref rc=$?"
timeout 60 build/fa_probe_tmem > gpurun_out/r2c3_probe_tmem.txt 2>&1; echo "probe rc=$?"
for v in fwd_base fwd_max3 fwd_ld fwd_p default; do
  if [ "$v" = default ]; then unset FA_SM100_LIB; else export FA_SM100_LIB=$PWD/build/variants/libfa_sm100_$v.so; fi
  timeout 120 python scripts/ab_time.py fwd >> gpurun_out/r2c3_ab_fwd.jsonl 2>> gpurun_out/r2c3_ab_fwd.err
done
unset FA_SM100_LIB
cat gpurun_out/r2c3_ab_fwd.jsonl
timeout 200 python tests/golden/make_golden_gpu.py > gpurun_out/r2c3_golden.log 2>&1; echo "golden rc=$?"
FA_BENCH_PREWARM_S=0 timeout 120 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c3_plain.log 2>&1 && \
FA_BENCH_PREWARM_S=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/r2c3_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c3_ncu1.log 2>&1
FA_BENCH_PREWARM_S=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -o gpurun_out/r2c3_prof_c4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c3_ncu2.log 2>&1
echo "ncu done rc=$?"
