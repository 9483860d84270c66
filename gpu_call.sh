mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/c4_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/c4_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/c4_pytest.log
timeout 600 python bench.py > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c4_bench.err; cat gpurun_out/c4_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/c4_bench_ref.json 2> gpurun_out/c4_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/c4_bench_ref.err; cat gpurun_out/c4_bench_ref.json
