"""Development check on a real B200: numerics vs an fp32 closed form computed on the GPU, plus
CUDA-event timings.  Not part of the product or of the pytest suite (tests/ holds those).

    python scripts/dev_check.py --what fwd --cases small,mid --time C2,C3
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch

import flashattn_b200 as fa
from flashattn_b200 import _cabi

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

CASES = {
    # name: (B, H, Sq, Sk, D, dtype, causal)
    "t1": (1, 1, 128, 128, 64, "bf16", False),
    "t2": (1, 1, 256, 256, 64, "bf16", False),
    "t3": (1, 2, 256, 256, 128, "bf16", False),
    "t4": (1, 2, 256, 256, 128, "bf16", True),
    "t5": (1, 2, 512, 512, 64, "fp16", True),
    "t6": (2, 3, 1024, 1024, 128, "fp16", True),
    "t7": (1, 2, 384, 640, 128, "bf16", False),      # Sq != Sk, Sq % 256 == 128
    "t8": (1, 2, 200, 300, 64, "bf16", False),       # ragged
    "t9": (1, 2, 333, 333, 128, "fp16", True),       # ragged causal
    "t10": (1, 1, 128, 4096, 128, "bf16", False),    # long K, one tile
    "c1": (1, 4, 512, 512, 64, "bf16", False),
    "c2": (4, 16, 2048, 2048, 64, "bf16", True),
    "c3": (4, 16, 4096, 4096, 128, "bf16", False),
    "c4s": (2, 32, 8192, 8192, 128, "bf16", True),
    "many": (8, 32, 512, 512, 128, "bf16", True),    # many items per CTA (scheduler wrap-around)
}
GROUPS = {
    "small": ["t1", "t2", "t3", "t4", "t5"],
    "mid": ["t6", "t7", "t8", "t9", "t10", "c1", "many"],
    "big": ["c2", "c3"],
}


def ref_closed_form(Q, K, V, dO, causal):
    """fp32 materialised attention on the GPU, per (b,h) slice to bound memory."""
    B, H, Sq, D = Q.shape
    Sk = K.shape[2]
    scale = 1.0 / math.sqrt(D)
    O = torch.empty(B, H, Sq, D, device=Q.device); LSE = torch.empty(B, H, Sq, device=Q.device)
    dQ = torch.empty_like(O); dK = torch.empty(B, H, Sk, D, device=Q.device); dV = torch.empty_like(dK)
    mask = None
    if causal:
        mask = ~(torch.arange(Sq, device=Q.device)[:, None] >= torch.arange(Sk, device=Q.device)[None, :])
    for b in range(B):
        q, k, v = Q[b].float(), K[b].float(), V[b].float()
        S = torch.matmul(q, k.transpose(-1, -2)) * scale
        if mask is not None:
            S.masked_fill_(mask, float("-inf"))
        lse = torch.logsumexp(S, dim=-1)
        P = torch.exp(S - lse[..., None])
        o = torch.matmul(P, v)
        O[b] = o; LSE[b] = lse
        if dO is not None:
            do = dO[b].float()
            dV[b] = torch.matmul(P.transpose(-1, -2), do)
            dP = torch.matmul(do, v.transpose(-1, -2))
            delta = (do * o).sum(-1, keepdim=True)
            dS = P * (dP - delta)
            dQ[b] = torch.matmul(dS, k) * scale
            dK[b] = torch.matmul(dS.transpose(-1, -2), q) * scale
    return O, LSE, dQ, dK, dV


def make(B, H, Sq, Sk, D, dt, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dtype = torch.bfloat16 if dt == "bf16" else torch.float16
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g).to(dtype)
    return mk(B, H, Sq, D), mk(B, H, Sk, D), mk(B, H, Sk, D), mk(B, H, Sq, D)


def err(a, b):
    a = a.float(); b = b.float()
    d = (a - b).abs()
    return d.max().item(), (d / (1e-2 + 1e-2 * b.abs())).max().item()


def run_case(name, what):
    B, H, Sq, Sk, D, dt, causal = CASES[name]
    Q, K, V, dO = make(B, H, Sq, Sk, D, dt)
    out = dict(case=name, shape=[B, H, Sq, Sk, D], dtype=dt, causal=causal)
    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
    torch.cuda.synchronize()
    rO, rLSE, rdQ, rdK, rdV = ref_closed_form(Q, K, V, dO if what == "all" else None, causal)
    out["O"] = err(O, rO); out["LSE"] = (LSE - rLSE).abs().max().item()
    ok = out["O"][1] < 1.0 and out["LSE"] < 1e-3 and bool(torch.isfinite(O.float()).all())
    if what == "all":
        dQ, dK, dV = fa.flash_attention_backward(Q, K, V, O, dO, LSE, causal)
        torch.cuda.synchronize()
        out["dQ"] = err(dQ, rdQ); out["dK"] = err(dK, rdK); out["dV"] = err(dV, rdV)
        ok = ok and out["dQ"][1] < 1.0 and out["dK"][1] < 1.0 and out["dV"][1] < 1.0
        # determinism: two runs bitwise equal
        dQ2, dK2, dV2 = fa.flash_attention_backward(Q, K, V, O, dO, LSE, causal)
        out["bwd_deterministic"] = bool(torch.equal(dQ, dQ2) and torch.equal(dK, dK2) and torch.equal(dV, dV2))
    O2, LSE2 = fa.flash_attention_forward(Q, K, V, causal)
    out["fwd_deterministic"] = bool(torch.equal(O, O2) and torch.equal(LSE, LSE2))
    out["ok"] = bool(ok)
    return out


def time_case(name, what, iters=20, warmup=5):
    B, H, Sq, Sk, D, dt, causal = CASES[name]
    Q, K, V, dO = make(B, H, Sq, Sk, D, dt)
    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def timeit(fn):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2], ts[0]
    res = dict(case=name)
    med, best = timeit(lambda: fa.flash_attention_forward(Q, K, V, causal))
    res["fwd_ms_med"] = med; res["fwd_ms_min"] = best
    res["fwd_tflops_med"] = fa.tflops(B, H, Sq, Sk, D, causal, "fwd", med)
    res["fwd_tflops_min"] = fa.tflops(B, H, Sq, Sk, D, causal, "fwd", best)
    if what == "all":
        med, best = timeit(lambda: fa.flash_attention_backward(Q, K, V, O, dO, LSE, causal))
        res["bwd_ms_med"] = med; res["bwd_ms_min"] = best
        res["bwd_tflops_med"] = fa.tflops(B, H, Sq, Sk, D, causal, "bwd", med)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="fwd", choices=["fwd", "all"])
    ap.add_argument("--cases", default="small")
    ap.add_argument("--time", default="")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    names = []
    for c in a.cases.split(","):
        if c:
            names += GROUPS.get(c, [c])
    lines = []
    rc = 0
    try:
        for n in names:
            r = run_case(n, a.what)
            lines.append(r); print(json.dumps(r), flush=True)
            if not r["ok"]:
                rc = 1
        for n in [x for x in a.time.split(",") if x]:
            r = time_case(n, a.what)
            lines.append(r); print(json.dumps(r), flush=True)
    except Exception as e:
        print("EXCEPTION", repr(e)[:500], flush=True)
        try:
            print("hang record:", _cabi.last_hang(), flush=True)
        except Exception as e2:
            print("no hang record:", repr(e2)[:200])
        rc = 2
    if a.out:
        with open(a.out, "a") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
    return rc


if __name__ == "__main__":
    sys.exit(main())
