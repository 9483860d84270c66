"""Multi-GPU configs of BASELINE.json (launch with torchrun, one rank per GPU):

  C4  B=16 H=32 N=8192 D=128 bf16 causal fwd+bwd, batch x head sharded (strong scaling: B/P per rank), no collective
  C5  B=1  H=32 N=131072 D=128 bf16 causal, sequence-sharded zigzag ring (K/V over NCCL P2P)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 scripts/bench_multi.py --config C5

Prints one JSON line per config on rank 0: ms (max over ranks, CUDA events), aggregate TFLOPS by the reference FLOP model
(code/Performance_Comparison.py:99-107), fraction of P x measured bf16 peak.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import flashattn_b200 as fa
import flashattn_b200.sharding as sh


def timeit(fn, iters, warmup, dev):
    for _ in range(warmup):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C4,C5")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--seq", type=int, default=131072)
    a = ap.parse_args()
    real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1590.0
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    mk = lambda *s: torch.randn(*s, device=dev, generator=g).bfloat16()
    for cfg in a.config.split(","):
        if cfg == "C4":
            B, H, N, D = 16, 32, 8192, 128
            if B % world:
                continue
            q, k, v, do = (mk(B // world, H, N, D) for _ in range(4))
            q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)

            def step():
                O = fa.flash_attention(q, k, v, True); O.backward(do)
                q.grad = None; k.grad = None; v.grad = None
            ms = timeit(step, a.iters, 2, dev)
            mode = "batch x head sharding, no collective"
        elif cfg == "C5":
            B, H, N, D = 1, 32, a.seq, 128
            if world < 2:
                continue
            q, k, v, do = (mk(B, H, N // world, D) for _ in range(4))       # this rank's zigzag shard (synthetic)
            q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)

            def step():
                O = sh.ring_flash_attention(q, k, v); O.backward(do)
                q.grad = None; k.grad = None; v.grad = None
            ms = timeit(step, a.iters, 1, dev)
            mode = "zigzag sequence-sharded ring, K/V + dK/dV over NCCL P2P"
        else:
            continue
        flops = 3.5 * 4 * B * H * N * N * D / 2
        tf = flops / (ms * 1e-3) / 1e12
        if rank == 0:
            real_stdout.write(json.dumps(dict(config=cfg, n_gpus=world, B=B, H=H, N=N, D=D, causal=True, dtype="bf16", mode=mode,
                                              ms_fwd_bwd=ms, tflops_aggregate=tf, tflops_per_gpu=tf / world,
                                              frac_of_measured_peak=tf / world / peak)) + "\n"); real_stdout.flush()
        del q, k, v, do
        torch.cuda.empty_cache()
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
