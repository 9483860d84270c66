"""The CPU oracle, pinned: against golden vectors produced by the reference's own Triton kernels
(tests/golden/make_golden.py) and against independent restatements of the same math."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(path):
    z = np.load(path)
    meta = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    arrs = {k: torch.from_numpy(z[k]).float() for k in z.files if not k.startswith("meta_")}
    return meta, arrs


def test_golden_fixtures_present():
    assert len(GOLD) >= 6


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_kernels(path):
    """Blocked restatement == reference Triton kernels: fp16 fixtures (unmodified kernels, CPU interpreter, 64x64 blocks) to
    <= 2.5 fp16 ulp; bf16 fixtures (bf16-patched kernels captured on a B200 with the autotuner's block sizes,
    tests/golden/make_golden_gpu.py) to <= 3 bf16 ulp."""
    m, g = _load(path)
    dt = getattr(torch, m["dtype"])
    Q, K, V, dO = orc.make_inputs(m["B"], m["H"], m["Sq"], m["Sk"], m["D"], dt, m["seed"])
    causal = bool(m["causal"])
    bm, bn = m.get("BLOCK_M", 64), m.get("BLOCK_N", 64)
    O, LSE = orc.forward_blocked(Q, K, V, causal, bm, bn)
    dQ, dK, dV, delta = orc.backward_blocked(Q, K, V, O, dO, LSE, causal, bm, bn)
    bf16 = dt == torch.bfloat16
    assert (LSE - g["LSE"]).abs().max() < (2e-5 if bf16 else 5e-6)
    for name, x in (("O", O), ("dQ", dQ), ("dK", dK), ("dV", dV)):
        ref = g[name]
        # <= 2.5 fp16 ulp (3 bf16 ulp) at the value's magnitude, with an absolute floor of 1 ulp at 0.5 for
        # near-zero sums (fp32 accumulation order differs between numpy / torch matmuls / the GPU's tile order)
        ulp = torch.clamp(ref.abs() * 2.0 ** (-7 if bf16 else -10), min=2.0 ** (-8 if bf16 else -11))
        assert ((x.float() - ref).abs() <= (3.0 if bf16 else 2.5) * ulp).all(), name
    if "delta" in g:
        assert (delta - g["delta"]).abs().max() < 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 128])
def test_blocked_vs_closed_form_and_sdpa(dtype, causal, D):
    Q, K, V, dO = orc.make_inputs(1, 2, 192, 192, D, dtype, seed=11)
    O, LSE = orc.forward_blocked(Q, K, V, causal)
    dQ, dK, dV, _ = orc.backward_blocked(Q, K, V, O, dO, LSE, causal)
    cO, cLSE, cdQ, cdK, cdV = orc.closed_form(Q, K, V, dO, causal)
    sO, sdQ, sdK, sdV = orc.sdpa_fp32(Q, K, V, dO, causal)
    # the two ground truths agree to fp32 round-off
    for a, b in ((cO, sO), (cdQ, sdQ), (cdK, sdK), (cdV, sdV)):
        assert (a.float() - b).abs().max() < 2e-5
    assert (LSE - cLSE.float()).abs().max() < 1e-3                    # Phase_3.md:752-753
    assert (orc.lse_bench(Q, K, causal) - cLSE.float()).abs().max() < 1e-5
    for a, b in ((O, cO), (dQ, cdQ), (dK, cdK), (dV, cdV)):            # BASELINE contract atol=rtol=1e-2
        assert torch.allclose(a.float(), b.float(), rtol=1e-2, atol=1e-2)


def test_block_size_independence():
    Q, K, V, dO = orc.make_inputs(1, 1, 256, 256, 64, torch.float16, seed=3)
    O1, L1 = orc.forward_blocked(Q, K, V, True, 64, 64)
    O2, L2 = orc.forward_blocked(Q, K, V, True, 128, 128)
    assert (L1 - L2).abs().max() < 1e-5
    assert (O1.float() - O2.float()).abs().max() < 2e-3


def test_cpu_flash_lse_matches():
    Q, K, V = orc.make_inputs(1, 2, 128, 128, 64, torch.float32, seed=5, with_dO=False)
    _, lse = orc.sdpa_cpu_flash(Q, K, V, None, True)
    assert (lse - orc.lse_bench(Q, K, True)).abs().max() < 1e-5


@pytest.mark.parametrize("causal", [False, True])
def test_merge_partials_is_attention_over_union(causal):
    """(O, LSE) merge over two disjoint key blocks == attention over all keys (ring algebra)."""
    Q, K, V = orc.make_inputs(1, 2, 128, 256, 64, torch.float32, seed=9, with_dO=False)
    O, LSE = orc.closed_form(Q, K, V, None, causal, q_offset=128 if causal else 0)
    Oa, La = orc.closed_form(Q, K[:, :, :128], V[:, :, :128], None, causal, q_offset=128 if causal else 0, k_offset=0)
    Ob, Lb = orc.closed_form(Q, K[:, :, 128:], V[:, :, 128:], None, causal, q_offset=128 if causal else 0, k_offset=128)
    Om, Lm = orc.merge_partials(Oa.float(), La.float(), Ob.float(), Lb.float())
    assert (Om - O.float()).abs().max() < 1e-5 and (Lm - LSE.float()).abs().max() < 1e-5
    # identity element
    ninf = torch.full_like(La.float(), float("-inf"))
    Oi, Li = orc.merge_partials(torch.zeros_like(Oa).float(), ninf, Oa.float(), La.float())
    assert torch.equal(Li, La.float()) and (Oi - Oa.float()).abs().max() < 1e-6


def test_range_mask_oracle_and_helpers():
    """closed_form(row_ranges=...) equals per-sequence attention for a packed batch, and the Ranges helpers are consistent:
    the key-side ranges describe the same mask as the query-side ranges."""
    import flashattn_b200 as fa
    cu = [0, 5, 12, 13, 20]
    total, H, D = cu[-1], 2, 16
    g = torch.Generator().manual_seed(3)
    Q, K, V, dO = (torch.randn(1, H, total, D, generator=g) for _ in range(4))
    r = fa.Ranges.from_cu_seqlens(cu, total)
    for causal in (False, True):
        O, LSE, dQ, dK, dV = orc.closed_form(Q, K, V, dO, causal, row_ranges=r.row_ranges())
        for s in range(len(cu) - 1):
            sl = slice(cu[s], cu[s + 1])
            o, lse, dq, dk, dv = orc.closed_form(Q[:, :, sl], K[:, :, sl], V[:, :, sl], dO[:, :, sl], causal)
            for a, b in ((O[:, :, sl], o), (LSE[:, :, sl], lse), (dQ[:, :, sl], dq), (dK[:, :, sl], dk), (dV[:, :, sl], dv)):
                assert (a - b).abs().max() < 1e-10
    for rr, Sq, Sk in ((r, total, total), (fa.Ranges.from_key_padding([3, 7], 6, 7), 6, 7), (fa.Ranges.sliding_window(2, 9, 4), 9, 9)):
        i = torch.arange(Sq)[None, :, None]; j = torch.arange(Sk)[None, None, :]
        from_rows = (j >= rr.row_lo[:, :, None]) & (j < rr.row_hi[:, :, None])
        from_cols = (i >= rr.col_lo[:, None, :]) & (i < rr.col_hi[:, None, :])
        if rr is not r and Sq == 9:                       # the window helper is meant to be combined with the causal mask
            from_rows &= (i >= j); from_cols &= (i >= j)
        assert torch.equal(from_rows, from_cols)
        for t in (rr.row_lo, rr.row_hi, rr.col_lo, rr.col_hi):
            assert (t[:, 1:] >= t[:, :-1]).all()         # monotone: the kernels take tile ranges from first / last rows


def test_dropout_mask_restatement():
    """dropout_keep_mask: keep rate = 1 - thresh/256, deterministic in the seed, and dropout in the closed form keeps the
    expectation (E[O_drop] = O) — the property the 1/(1-p) rescale exists for."""
    k1, s1 = orc.dropout_keep_mask(7, 2, 3, 96, 160, 0.25)
    k2, _ = orc.dropout_keep_mask(7, 2, 3, 96, 160, 0.25)
    k3, _ = orc.dropout_keep_mask(8, 2, 3, 96, 160, 0.25)
    assert torch.equal(k1, k2) and not torch.equal(k1, k3) and abs(s1 - 4 / 3) < 1e-12
    assert abs(k1.float().mean().item() - 0.75) < 5e-3
    assert (k1.float().mean(dim=(0, 1, 2)) - 0.75).abs().max() < 0.08          # no dead key columns
    assert orc.dropout_keep_mask(7, 1, 1, 8, 8, 0.0)[0].all()
    g = torch.Generator().manual_seed(5)
    Q, K, V = (torch.randn(1, 1, 32, 16, generator=g) for _ in range(3))
    O = orc.closed_form(Q, K, V)[0]
    acc = torch.zeros_like(O)
    n = 400
    for seed in range(n):
        keep, sc = orc.dropout_keep_mask(seed, 1, 1, 32, 32, 0.5)
        acc += orc.closed_form(Q, K, V, keep_mask=keep, keep_scale=sc)[0]
    assert (acc / n - O).abs().max() < 0.12
