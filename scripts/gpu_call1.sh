#!/bin/bash
# round 2, GPU call 1: smoke, full GPU test suite, headline bench (both arms), bf16 goldens, ncu launch list + full capture on C4
mkdir -p gpurun_out && rm -f gpurun_out/parity_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_env.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/r2c1_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q -rfEs --durations=15 > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2c1_pytest.log
python bench.py > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/r2c1_bench_ref.json 2> gpurun_out/r2c1_bench_ref.err; echo "bench ref rc=$?"
python tests/golden/make_golden_gpu.py > gpurun_out/r2c1_golden.log 2>&1; echo "golden rc=$?"
FA_BENCH_PREWARM_S=0 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c1_plain.log 2>&1 && \
FA_BENCH_PREWARM_S=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 24 --csv --log-file gpurun_out/r2c1_launches_c4.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c1_ncu1.log 2>&1
FA_BENCH_PREWARM_S=0 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c1_plain2.log 2>&1 && \
FA_BENCH_PREWARM_S=0 ncu --set full --clock-control none --import-source on -k regex:fa_ -s 12 -c 4 -o gpurun_out/r2c1_prof_c4 python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r2c1_ncu2.log 2>&1
echo "ncu done rc=$?"
