"""Generate golden vectors by running the REFERENCE'S OWN Triton kernels on the CPU.

Run in the build container only (needs /root/reference and triton's interpreter):

    TRITON_INTERPRET=1 python tests/golden/make_golden.py

The three kernels in /root/reference/code/_flash_attention_kernel_optimized.py are called
through ``kernel.fn[grid]`` with BLOCK_M = BLOCK_N = 64 (the Autotuner wrapper needs a GPU
driver; ``.fn`` does not — SURVEY App. B.1) using the same flattened TensorDescriptors the
reference launcher builds (code/My_FlashAttention_optimized.py:33-51, :79-108).

Only fp16 cases are generated: they run the reference UNMODIFIED.  bf16 cannot be pinned
this way — the shipped kernels assert on bf16 and, with the casts retargeted,
the interpreter's numpy-backed ``.to(tl.bfloat16)`` does not round to nearest-even (probed:
up to 1 ulp off torch's cast), so its bf16 outputs are not the reference's arithmetic.
bf16 reference outputs are captured on a real B200 instead (``tests/golden/make_golden_gpu.py``).

Inputs are NOT stored: they are regenerated from ``oracle.attention_oracle.make_inputs``
with the recorded seed.  Outputs (O, LSE, dQ, dK, dV, delta) are stored as fp32/fp16 arrays.
"""
import os
import sys

os.environ["TRITON_INTERPRET"] = "1"
REF = "/root/reference/code"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))

import numpy as np
import torch
from triton.tools.tensor_descriptor import TensorDescriptor as TD

from oracle.attention_oracle import make_inputs

BM = BN = 64

CASES = [
    # name,            B, H, Sq,  Sk,  D,   dtype,     causal, seed
    ("fp16_d64_nc",    1, 2, 128, 128, 64,  "float16",  False, 1),
    ("fp16_d64_c",     1, 2, 128, 128, 64,  "float16",  True,  2),
    ("fp16_d128_nc",   1, 2, 128, 128, 128, "float16",  False, 3),
    ("fp16_d128_c",    1, 2, 256, 256, 128, "float16",  True,  4),
    ("fp16_d64_cross", 1, 2, 128, 256, 64,  "float16",  False, 5),
    ("fp16_selfcheck", 1, 2, 256, 256, 64,  "float16",  True,  42),   # the reference __main__ shape, shrunk in B,H
]


def load_kernels():
    import _flash_attention_kernel_optimized as kk       # the unmodified reference module
    return kk


def run_case(B, H, Sq, Sk, D, dtype, causal, seed):
    dt = getattr(torch, dtype)
    kk = load_kernels()
    Q, K, V, dO = make_inputs(B, H, Sq, Sk, D, dt, seed)
    O = torch.empty_like(Q); LSE = torch.empty(B, H, Sq)
    dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V)
    delta = torch.empty(B, H, Sq)
    d2 = lambda t, S, rows: TD(t, [B * H * S, D], [D, 1], [rows, D], "zero")
    d1 = lambda t, S, rows: TD(t, [B * H * S], [1], [rows], "zero")
    scale = 1 / (D ** 0.5)
    kk.flash_attention_forward_kernel.fn[(Sq // BM, B * H)](
        d2(Q, Sq, BM), d2(K, Sk, BN), d2(V, Sk, BN), d2(O, Sq, BM), d1(LSE, Sq, BM),
        scale, B, H, Sq, Sk, D, BLOCK_M=BM, BLOCK_N=BN, is_causal=causal)
    kk.flash_attention_dQ_kernel.fn[(Sq // BM, B * H)](
        d2(Q, Sq, BM), d2(K, Sk, BN), d2(V, Sk, BN), d2(dO, Sq, BM), d2(O, Sq, BM), d1(LSE, Sq, BM),
        d2(dQ, Sq, BM), d1(delta, Sq, BM),
        scale, B, H, Sq, Sk, D, BLOCK_M=BM, BLOCK_N=BN, is_causal=causal)
    kk.flash_attention_dKV_kernel.fn[(Sk // BN, B * H)](
        d2(Q, Sq, BM), d2(K, Sk, BN), d2(V, Sk, BN), d2(dO, Sq, BM), d1(LSE, Sq, BM),
        d2(dK, Sk, BN), d2(dV, Sk, BN), d1(delta, Sq, BM),
        scale, B, H, Sq, Sk, D, BLOCK_M=BM, BLOCK_N=BN, is_causal=causal)
    return dict(O=O, LSE=LSE, dQ=dQ, dK=dK, dV=dV, delta=delta)


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, B, H, Sq, Sk, D, dtype, causal, seed in CASES:
        res = run_case(B, H, Sq, Sk, D, dtype, causal, seed)
        arrs = {}
        for k, v in res.items():
            arrs[k] = v.numpy()                      # fp16 stays fp16 (exact), LSE/delta fp32
        meta = dict(B=B, H=H, Sq=Sq, Sk=Sk, D=D, dtype=dtype, causal=int(causal), seed=seed,
                    BLOCK_M=BM, BLOCK_N=BN, patched=0)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"),
                            **arrs, **{f"meta_{k}": np.array(v) for k, v in meta.items()})
        print(name, {k: float(np.abs(a.astype(np.float32)).max()) for k, a in arrs.items()})


if __name__ == "__main__":
    main()
