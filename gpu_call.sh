mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "out_of_bounds" 2>&1 | tail -3
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/c21_bench$N.json 2> gpurun_out/c21_bench$N.err; echo "bench$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c21_bench$N.json')); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N scripts/bench_multi.py --config C4,C5 > gpurun_out/c21_multi$N.jsonl 2> gpurun_out/c21_multi$N.err; echo "multi$N rc=$?"; cat gpurun_out/c21_multi$N.jsonl | cut -c1-420
done
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
