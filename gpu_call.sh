mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke > gpurun_out/c18_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c18_smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c18_pytest.log
timeout 600 python bench.py > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c18_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/c18_bench_ref.json 2> gpurun_out/c18_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload C3 > gpurun_out/c18_bench_c3.json 2> gpurun_out/c18_bench_c3.err; echo "c3 rc=$?"
timeout 600 python bench.py --workload C3 --impl reference > gpurun_out/c18_bench_c3_ref.json 2> /dev/null; echo "c3 ref rc=$?"
python - <<'PY'
import json
for f in ("c18_bench","c18_bench_ref","c18_bench_c3","c18_bench_c3_ref"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("clocks"))
    if "kernels" in d: print({k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d["kernels"].items()}); print({k:(round(v["fwd_tflops"]), round(v["fwd_bwd_tflops"])) for k,v in d.get("also",{}).items()})
PY
