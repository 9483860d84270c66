"""Bring-up of the fused backward: compare against the two-kernel path on the same inputs, then time both."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
from flashattn_b200 import interface as I, _cabi

torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(B, H, Hk, Sq, Sk, causal, dtype=torch.bfloat16, D=64):
    Q = torch.randn(B, H, Sq, D, device="cuda").to(dtype); dO = torch.randn(B, H, Sq, D, device="cuda").to(dtype)
    K = torch.randn(B, Hk, Sk, D, device="cuda").to(dtype); V = torch.randn(B, Hk, Sk, D, device="cuda").to(dtype)
    O, LSE = I.flash_attention_forward(Q, K, V, causal)
    outs = []
    for fused in (False, True):
        dQ, dK, dV = torch.full_like(Q, float("nan")), torch.full_like(K, float("nan")), torch.full_like(V, float("nan"))
        delta = torch.empty(B, H, Sq, device="cuda")
        if fused:
            I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal)
        else:
            I.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, 7)
        torch.cuda.synchronize()
        outs.append((dQ.float(), dK.float(), dV.float()))
    hang = _cabi.last_hang()
    res = {"shape": [B, H, Hk, Sq, Sk], "causal": causal, "dtype": str(dtype), "hang": hang}
    for name, a, b in zip(("dQ", "dK", "dV"), outs[0], outs[1]):
        res[name] = {"maxdiff": float((a - b).abs().max()), "ref_absmax": float(a.abs().max()), "nan": bool(torch.isnan(b).any()),
                     "equal": bool(torch.equal(a, b))}
    return res


def timeit(fn, iters=12, warmup=3):
    for _ in range(warmup): fn()
    ts = []
    for _ in range(iters):
        flush.zero_(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return round(ts[len(ts) // 2], 4)


def bench(B, H, S, causal, D=64):
    Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda").bfloat16() for _ in range(4))
    O, LSE = I.flash_attention_forward(Q, K, V, causal)
    dQ, dK, dV = torch.empty_like(Q), torch.empty_like(K), torch.empty_like(V)
    delta = torch.empty(B, H, S, device="cuda"); acc = torch.empty(B, H, S, D, device="cuda")
    t2 = timeit(lambda: I.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, 7))
    tf = timeit(lambda: I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, dq_acc=acc))
    parts = {n: timeit(lambda: I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, dq_acc=acc, parts=m))
             for n, m in (("delta0", 1), ("fused", 8), ("convert", 16))}
    fl = 2.5 * 4 * B * H * S * S * D / (2 if causal else 1)
    return {"bench": [B, H, S, D], "causal": causal, "two_kernel_ms": t2, "fused_ms": tf, "parts": parts,
            "two_kernel_tflops": round(fl / t2 / 1e9, 1), "fused_tflops": round(fl / tf / 1e9, 1)}


SHAPES128 = [(1, 1, 1, 128, 128, False), (1, 2, 2, 256, 256, True), (2, 4, 4, 512, 512, False), (1, 2, 2, 200, 300, False),
             (1, 3, 3, 333, 333, True), (2, 8, 2, 320, 448, False), (2, 8, 2, 320, 320, True), (1, 2, 2, 1, 77, False),
             (2, 4, 4, 2048, 2048, True), (1, 2, 2, 8192, 8192, True), (1, 160, 160, 1024, 1024, True)]

if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "check128":          # one subprocess per shape: a trapped kernel poisons only its own CUDA context
        import subprocess
        for i in range(len(SHAPES128) + 1):
            r = subprocess.run([sys.executable, __file__, "one128", str(i)], capture_output=True, text=True, timeout=180)
            print(r.stdout.strip() or json.dumps({"shape_index": i, "rc": r.returncode, "stderr": r.stderr[-600:]}), flush=True)
        sys.exit(0)
    if what == "one128":
        i = int(sys.argv[2])
        if i < len(SHAPES128):
            print(json.dumps(run(*SHAPES128[i], D=128)), flush=True)
        else:
            print(json.dumps(run(2, 4, 4, 512, 512, True, torch.float16, D=128)), flush=True)
        sys.exit(0)
    if what == "bench128":
        for args in [(4, 16, 4096, False), (2, 32, 8192, True), (4, 8, 1024, True), (16, 32, 8192, True)]:
            print(json.dumps(bench(*args, D=128)), flush=True)
        sys.exit(0)
    if what in ("all", "check"):
        for args in [(1, 1, 1, 128, 128, False), (1, 2, 2, 256, 256, True), (2, 4, 4, 512, 512, False), (1, 2, 2, 200, 300, False),
                     (1, 3, 3, 333, 333, True), (2, 8, 2, 320, 448, False), (2, 8, 2, 320, 320, True), (1, 2, 2, 1, 77, False),
                     (4, 16, 16, 2048, 2048, True)]:
            print(json.dumps(run(*args)), flush=True)
        print(json.dumps(run(2, 4, 4, 512, 512, True, torch.float16)), flush=True)
    if what in ("all", "bench"):
        for args in [(4, 16, 2048, True), (4, 16, 2048, False), (1, 16, 8192, True), (1, 16, 8192, False), (8, 16, 512, True)]:
            print(json.dumps(bench(*args)), flush=True)
