import torch

import flashattn_b200 as fa


def test_verify_results_metrics_and_verdict():
    b = torch.tensor([1.0, 2.0, -3.0, 0.0])
    r = fa.verify_results(b, b.clone())
    assert r["passed"] and r["max_abs_err"] == 0 and abs(r["cosine_sim"] - 1) < 1e-12
    t = b + torch.tensor([0.0, 0.05, 0.0, 0.0])        # 0.05 > atol + rtol*|t| = 1e-3 + 2.05e-2
    r = fa.verify_results(b, t)
    assert not r["allclose"] and not r["passed"]
    assert abs(r["max_abs_err"] - 0.05) < 1e-6
    assert abs(r["max_norm_err"] - 0.05 / (1e-3 + 1e-2 * 2.05)) < 1e-3
    r = fa.verify_results(b, t, rtol=1e-1, atol=1e-1)
    assert r["passed"]


def test_flop_model_matches_reference_formula():
    # code/Performance_Comparison.py:101-107 ; BASELINE.md §3 numbers
    assert fa.attention_flops(4, 16, 2048, 2048, 64, True, "fwd") == 34359738368
    assert fa.attention_flops(4, 16, 4096, 4096, 128, False, "fwd") == 549755813888
    assert fa.attention_flops(4, 16, 2048, 2048, 64, True, "fwd_bwd") == 3.5 * 34359738368
    assert fa.attention_flops(1, 4, 512, 512, 64, False, "bwd") == 2.5 * 268435456
    assert abs(fa.tflops(4, 16, 4096, 4096, 128, False, "fwd", 0.5) - 1099.511627776) < 1e-6
