"""dKV/dQ-only A/B through the raw C ABI (works across ABI revisions of the strided entry points)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
lib = ctypes.CDLL(os.environ["FA_SM100_LIB"])
vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
lib.fa_sm100_fwd.argtypes = [vp]*5 + [i]*7 + [f, vp]; lib.fa_sm100_fwd.restype = i
lib.fa_sm100_bwd_parts.argtypes = [vp]*10 + [i]*7 + [f, vp, i]; lib.fa_sm100_bwd_parts.restype = i
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, iters=12, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return round(ts[len(ts)//2], 4)
out = {"lib": os.path.basename(os.environ["FA_SM100_LIB"])}
st = torch.cuda.current_stream().cuda_stream
for B, H, S, D, c in [(4,16,2048,64,1), (1,16,8192,64,0), (4,16,4096,128,0)]:
    Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda").bfloat16() for _ in range(4))
    O = torch.empty_like(Q); LSE = torch.empty(B, H, S, device="cuda"); delta = torch.empty_like(LSE)
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    p = lambda t: t.data_ptr()
    assert lib.fa_sm100_fwd(p(Q), p(K), p(V), p(O), p(LSE), B, H, S, S, D, 1, c, 0.0, st) == 0
    run = lambda parts: lib.fa_sm100_bwd_parts(p(Q), p(K), p(V), p(O), p(dO), p(LSE), p(dQ), p(dK), p(dV), p(delta), B, H, S, S, D, 1, c, 0.0, st, parts)
    assert run(7) == 0
    out[f"{S}x{D}{'c' if c else 'n'}"] = {"dQ": timeit(lambda: run(2)), "dKV": timeit(lambda: run(4))}
print(json.dumps(out))
