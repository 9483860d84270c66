"""``verify_results`` of the reference (code/_verify_func.py:3-40), returning the metrics
instead of printing them.  Same five metrics, same pass rule: allclose(rtol, atol) and cosine > 0.999."""
from __future__ import annotations

import torch


def verify_results(bench, output, name="Attention", rtol=1e-2, atol=1e-3, verbose=False):
    b = bench.detach().to(torch.float32)                     # :5
    t = output.detach().to(torch.float32)                    # :6
    diff_abs = (b - t).abs()                                 # :7
    max_abs_err = diff_abs.max().item()                      # :10
    mean_abs_err = diff_abs.mean().item()                    # :11
    max_rel_err = (diff_abs / (b.abs() + 1e-5)).max().item() # :14-15
    max_norm = (diff_abs / (atol + rtol * t.abs())).max().item()   # :18-20, < 1 passes
    cosine_sim = torch.nn.functional.cosine_similarity(b.flatten().double(), t.flatten().double(), dim=0).item()  # :23-25
    is_allclose = torch.allclose(b, t, rtol=rtol, atol=atol) # :35
    passed = bool(is_allclose and cosine_sim > 0.999)        # :37
    res = dict(name=name, max_abs_err=max_abs_err, mean_abs_err=mean_abs_err, max_rel_err=max_rel_err,
               max_norm_err=max_norm, cosine_sim=cosine_sim, allclose=bool(is_allclose), passed=passed)
    if verbose:
        print(f"[{name}] max_abs={max_abs_err:.2e} mean_abs={mean_abs_err:.2e} max_rel={max_rel_err:.2e} "
              f"max_norm={max_norm:.2e} cos={cosine_sim:.6f} {'PASS' if passed else 'FAIL'}")
    return res
