"""Capture bf16 reference outputs on a real B200: the reference's Triton kernels with their four hard ``tl.float16``
dot-operand casts retargeted to ``tl.bfloat16`` (the shipped kernels assert on bf16 inputs; the patch is the one
``baseline/ref_runner.py`` applies, nothing else changes), run through the reference's own ``flash_attention()``.

Run on the GPU box (the copy of the reference in baseline/_ref travels with the repo snapshot):

    python tests/golden/make_golden_gpu.py            # writes gpurun_out/golden/bf16_*.npz

then copy the files into tests/golden/ in the build container and commit them.  The fp16 fixtures next to them come from the
UNMODIFIED kernels through the Triton CPU interpreter (make_golden.py); the interpreter's numpy bf16 cast is not
round-to-nearest-even, which is why the bf16 ones are captured on the GPU instead.

Inputs are NOT stored: they are regenerated from ``oracle.attention_oracle.make_inputs`` with the recorded seed.  bf16 outputs
are stored as fp32 (exact superset), meta_dtype = "bfloat16", meta_patched = 1.
"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline"))

import numpy as np
import torch

import ref_runner
from oracle.attention_oracle import make_inputs

CASES = [
    # name,             B, H, Sq,  Sk,  D,   causal, seed
    ("bf16_d64_c",      1, 2, 256, 256, 64,  True,  12),
    ("bf16_d64_nc",     1, 2, 128, 256, 64,  False, 13),
    ("bf16_d128_c",     1, 2, 256, 256, 128, True,  14),
    ("bf16_d128_nc",    1, 2, 256, 256, 128, False, 15),
]


def main():
    why = ref_runner.available()
    if why:
        raise SystemExit(why)
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    mod = ref_runner._load(True)
    for name, B, H, Sq, Sk, D, causal, seed in CASES:
        Q, K, V, dO = make_inputs(B, H, Sq, Sk, D, torch.bfloat16, seed)
        q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
        O = mod.flash_attention(q, k, v, causal)
        O.backward(dO.cuda())
        _, LSE = mod.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal)
        torch.cuda.synchronize()
        arrs = {n: t.detach().float().cpu().numpy() for n, t in (("O", O), ("LSE", LSE), ("dQ", q.grad), ("dK", k.grad), ("dV", v.grad))}
        meta = dict(B=B, H=H, Sq=Sq, Sk=Sk, D=D, dtype="bfloat16", causal=int(causal), seed=seed, patched=1)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), **arrs, **{f"meta_{k}": np.array(v) for k, v in meta.items()})
        print(name, {k: float(np.abs(a).max()) for k, a in arrs.items()})


if __name__ == "__main__":
    main()
