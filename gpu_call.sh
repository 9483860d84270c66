mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N > gpurun_out/c31_bench$N.json 2> gpurun_out/c31_bench$N.err; echo "bench$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c31_bench$N.json')); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/bench_multi.py --config C4,C5 > gpurun_out/c31_multi$N.jsonl 2> gpurun_out/c31_multi$N.err; echo "multi$N rc=$?"; cut -c1-330 gpurun_out/c31_multi$N.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --impl reference > gpurun_out/c31_bench${N}_ref.json 2> /dev/null; echo "ref$N rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c31_bench${N}_ref.json')); print('ref', d['n_gpus'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
