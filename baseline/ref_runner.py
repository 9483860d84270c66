"""Run the UNMODIFIED reference (its Triton kernels, through its own public API) on this GPU.

``baseline/_ref/`` is a git-ignored copy of /root/reference/code/*.py made by
``__graft_entry__.build()`` in the build container (the reference has no setup.py, so
``pip install`` cannot be used — recorded in DESIGN.md).  It travels to the GPU box with the
repo snapshot.  Nothing in the product imports this module; ``bench.py --impl reference`` and
the comparison tests do.

* fp16: the reference exactly as shipped (code/My_FlashAttention_optimized.py:169-170).
* bf16: the shipped kernels assert (`Both operands must be same dtype`).  A
  second copy with the hard ``tl.float16`` casts retargeted to ``tl.bfloat16`` is generated
  at run time into a temp dir and is always labelled ``"patched": true``.
Timing follows code/Performance_Comparison.py:111-128 (warm-up 10, repeat 30, one CUDA-event
pair around the repeat loop, through the autograd entry) and the FLOP model of :99-107.
"""
from __future__ import annotations

import importlib
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("FA_REF_DIR", os.path.join(HERE, "_ref"))


def available() -> str | None:
    """None if the reference copy can be used here, else a one-line reason."""
    if not os.path.isfile(os.path.join(REF_DIR, "My_FlashAttention_optimized.py")):
        return f"no reference copy under {REF_DIR} (run __graft_entry__.build() where /root/reference exists)"
    try:
        import torch
        if not torch.cuda.is_available():
            return "no CUDA device (the reference asserts Q.is_cuda)"
        import triton  # noqa: F401
    except Exception as e:  # pragma: no cover
        return f"import failed: {e!r}"
    return None


def _load(patched_bf16: bool):
    """Import the reference operator module (optionally the bf16-patched copy)."""
    if not patched_bf16:
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        return importlib.import_module("My_FlashAttention_optimized")
    d = os.path.join(tempfile.gettempdir(), "fa_ref_bf16_patched")
    os.makedirs(d, exist_ok=True)
    ksrc = open(os.path.join(REF_DIR, "_flash_attention_kernel_optimized.py")).read()
    with open(os.path.join(d, "_flash_attention_kernel_optimized_bf16.py"), "w") as f:
        f.write(ksrc.replace("tl.float16", "tl.bfloat16"))
    osrc = open(os.path.join(REF_DIR, "My_FlashAttention_optimized.py")).read()
    with open(os.path.join(d, "My_FlashAttention_optimized_bf16.py"), "w") as f:
        f.write(osrc.replace("from _flash_attention_kernel_optimized import",
                             "from _flash_attention_kernel_optimized_bf16 import"))
    if d not in sys.path:
        sys.path.insert(0, d)
    return importlib.import_module("My_FlashAttention_optimized_bf16")


def ref_flash_attention(dtype_is_bf16: bool):
    return _load(dtype_is_bf16).flash_attention


def flops_fwd(B, H, Sq, Sk, D, causal):
    """code/Performance_Comparison.py:101"""
    return 4 * B * H * Sq * Sk * D // (2 if causal else 1)


def timing(run_fn, warmup=10, repeat=30):
    """code/Performance_Comparison.py:111-128"""
    import torch
    for _ in range(warmup):
        run_fn()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(repeat):
        run_fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / repeat


def bench_fn(fn, B, H, Sq, Sk, D, causal, dtype, warmup=10, repeat=30, seed=0):
    """Time `fn(Q,K,V,is_causal)` fwd and fwd+bwd the way the reference benchmark does
    (code/Performance_Comparison.py:38-93).  Returns dict of ms and TFLOPS."""
    import torch
    torch.manual_seed(seed)
    dev = torch.device("cuda", torch.cuda.current_device())
    Q = torch.randn(B, H, Sq, D, device=dev, dtype=torch.float32).to(dtype).requires_grad_(True)
    K = torch.randn(B, H, Sk, D, device=dev, dtype=torch.float32).to(dtype).requires_grad_(True)
    V = torch.randn(B, H, Sk, D, device=dev, dtype=torch.float32).to(dtype).requires_grad_(True)
    dO = torch.randn(B, H, Sq, D, device=dev, dtype=torch.float32).to(dtype)

    def run_fwd():
        return fn(Q, K, V, causal)

    def run_all():
        O = fn(Q, K, V, causal)
        O.backward(dO)
        Q.grad = None; K.grad = None; V.grad = None
        return O

    t_fwd = timing(run_fwd, warmup, repeat)
    t_all = timing(run_all, warmup, repeat)
    f = flops_fwd(B, H, Sq, Sk, D, causal)
    t_bwd = max(t_all - t_fwd, 1e-9)
    return dict(ms_fwd=t_fwd, ms_fwd_bwd=t_all, ms_bwd=t_bwd,
                tflops_fwd=f / (t_fwd * 1e-3) / 1e12,
                tflops_bwd=2.5 * f / (t_bwd * 1e-3) / 1e12,
                tflops_fwd_bwd=3.5 * f / (t_all * 1e-3) / 1e12)


def sdpa_flash(dtype):
    """The reference's own yardstick (code/Performance_Comparison.py:53-57)."""
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel

    def fn(Q, K, V, causal):
        with sdpa_kernel(SDPBackend.FLASH_ATTENTION):
            return F.scaled_dot_product_attention(Q, K, V, is_causal=causal)
    return fn


CONFIGS = {
    "C2": dict(B=4, H=16, Sq=2048, Sk=2048, D=64, causal=True),
    "C3": dict(B=4, H=16, Sq=4096, Sk=4096, D=128, causal=False),
    "C4s": dict(B=2, H=32, Sq=8192, Sk=8192, D=128, causal=True),   # one 8-GPU shard of C4 (B=16/8)
    "sweep4k": dict(B=4, H=8, Sq=4096, Sk=4096, D=128, causal=True),  # code/Performance_Comparison.py:152-162
}


def main(argv=None):
    import argparse
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3")
    ap.add_argument("--dtypes", default="fp16,bf16")
    ap.add_argument("--sdpa", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args(argv)
    why = available()
    if why:
        print(json.dumps({"impl": "reference", "unavailable": why})); return 0
    lines = []
    for cname in a.configs.split(","):
        cfg = CONFIGS[cname]
        for dn in a.dtypes.split(","):
            dt = torch.bfloat16 if dn == "bf16" else torch.float16
            try:
                r = bench_fn(ref_flash_attention(dn == "bf16"), dtype=dt, **cfg)
                line = dict(impl="reference-triton", config=cname, dtype=dn, patched=(dn == "bf16"), **cfg, **r)
            except Exception as e:
                line = dict(impl="reference-triton", config=cname, dtype=dn, error=repr(e)[:300])
            lines.append(line); print(json.dumps(line), flush=True)
            if a.sdpa:
                r = bench_fn(sdpa_flash(dt), dtype=dt, **cfg)
                line = dict(impl="torch-sdpa-flash", config=cname, dtype=dn, **cfg, **r)
                lines.append(line); print(json.dumps(line), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            for l in lines:
                f.write(json.dumps(l) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
