"""A/B timing of one library build (FA_SM100_LIB) over a few shapes: prints fwd / dQ / dKV ms and errors.
   python scripts/ab_time.py [fwd|all]"""
import os, sys, json, math
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
shapes = [(4, 16, 2048, 64, 1), (1, 16, 8192, 64, 0), (4, 16, 4096, 128, 0), (2, 32, 8192, 128, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=12, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
out = {"lib": os.path.basename(os.environ.get("FA_SM100_LIB", "default"))}
for B, H, S, D, c in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(4))
    O, LSE = fa.flash_attention_forward(Q, K, V, bool(c))
    key = f"{S}x{D}{'c' if c else 'n'}"
    r = {"fwd": round(timeit(lambda: fa.flash_attention_forward(Q, K, V, bool(c))), 4)}
    if what == "all":
        dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V); delta = torch.empty(B, H, S, device="cuda")
        fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, bool(c), 7)
        r["dQ"] = round(timeit(lambda: fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, bool(c), 2)), 4)
        r["dKV"] = round(timeit(lambda: fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, bool(c), 4)), 4)
    out[key] = r
# accuracy on one (b,h) slice vs fp32
torch.backends.cuda.matmul.allow_tf32 = False
Q, K, V = (torch.randn(1, 2, 1024, 128, device="cuda").bfloat16() for _ in range(3))
O, LSE = fa.flash_attention_forward(Q, K, V, True)
S_ = (Q.float() @ K.float().transpose(-1, -2)) / math.sqrt(128)
i = torch.arange(1024, device="cuda"); S_.masked_fill_(~(i[:, None] >= i[None, :]), float("-inf"))
lse = torch.logsumexp(S_, -1); ref = torch.softmax(S_, -1) @ V.float()
out["err_O_max"] = round((O.float() - ref).abs().max().item(), 5); out["err_O_mean"] = float(f"{(O.float() - ref).abs().mean().item():.3e}")
out["err_LSE"] = float(f"{(LSE - lse).abs().max().item():.2e}")
print(json.dumps(out), flush=True)
