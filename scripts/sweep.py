"""Per-kernel timing sweep (development aid): ms, algorithmic TFLOP/s and SM-clocks per 128x128 score tile
per SM for fwd / dQ / dKV over a grid of shapes.  python scripts/sweep.py [--clock-ghz 1.9]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="4,16,2048,64,1;4,16,2048,64,0;1,16,8192,64,1;1,16,8192,64,0;4,16,2048,128,1;4,16,4096,128,0;2,32,8192,128,1;8,16,512,64,1;8,16,1024,128,1")
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=a.iters, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]

for sh in a.shapes.split(";"):
    B, H, S, D, causal = map(int, sh.split(",")); causal = bool(causal)
    g = torch.Generator(device="cuda").manual_seed(0)
    Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(4))
    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
    dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V)
    delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, 7)
    t = {"fwd": timeit(lambda: fa.flash_attention_forward(Q, K, V, causal))}
    for name, part in (("delta", 1), ("dQ", 2), ("dKV", 4)):
        t[name] = timeit(lambda: fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, part))
    nt = S // 128
    tiles = B * H * (nt * (nt + 1) // 2 if causal else nt * nt)          # 128x128 score tiles
    gemm = 2.0 * B * H * S * S * D / (2 if causal else 1)
    out = dict(shape=[B, H, S, D], causal=causal, tiles_per_sm=round(tiles / 148, 1))
    for k, n in (("fwd", 2), ("dQ", 3), ("dKV", 4)):
        out[k] = dict(ms=round(t[k], 4), tflops=round(n * gemm / t[k] / 1e9, 1), us_per_tile_per_sm=round(t[k] * 1e3 / (tiles / 148), 3))
    out["delta_ms"] = round(t["delta"], 4)
    tot = sum(t.values())
    out["fwd_bwd_tflops"] = round(3.5 * 2 * gemm / tot / 1e9, 1)
    print(json.dumps(out), flush=True)
