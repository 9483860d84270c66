"""A/B timing of the fused D=128 backward kernel alone (parts = FUSED) for one library build (FA_SM100_LIB)."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from flashattn_b200 import interface as I
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return round(ts[len(ts) // 2], 4)
out = {"lib": os.path.basename(os.environ.get("FA_SM100_LIB", "default"))}
for B, H, S, c in ((4, 16, 4096, 0), (2, 32, 8192, 1)):
    g = torch.Generator(device="cuda").manual_seed(0)
    Q, K, V, dO = (torch.randn(B, H, S, 128, device="cuda", generator=g).bfloat16() for _ in range(4))
    O, LSE = I.flash_attention_forward(Q, K, V, bool(c))
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    delta = torch.empty(B, H, S, device="cuda"); acc = torch.zeros(B, H, S, 128, device="cuda")
    I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, bool(c), dq_acc=acc, parts=1)
    t = timeit(lambda: I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, bool(c), dq_acc=acc, parts=8))
    pairs = B * H * (S // 128) * ((S // 128 + 1) / 2 if c else S // 128)
    out[f"{S}{'c' if c else 'n'}"] = {"fused_ms": t, "us_per_pair_per_sm": round(t * 1e3 / (pairs / 148), 3)}
print(json.dumps(out), flush=True)
