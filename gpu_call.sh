TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR bench.py --gpus 8 --steps 30 --warmup 10 > gpurun_out/c62_bench8.json 2> gpurun_out/c62_bench8.err; tail -1 gpurun_out/c62_bench8.err | cut -c1-200
timeout 200 $TR scripts/bench_multi.py --config C4,C5 > gpurun_out/c62_multi8.jsonl 2> gpurun_out/c62_multi8.err
cat gpurun_out/c62_multi8.jsonl | cut -c1-700
python - <<'PY'
import json
d = json.load(open("gpurun_out/c62_bench8.json")); print(d["value"], d["ms_per_step"], d["n_gpus"], d["e2e"]["value"], d["clocks"])
PY
