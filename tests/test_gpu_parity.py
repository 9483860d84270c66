"""Parity of the sm_100a CUDA path (through the C ABI) against the CPU oracle, the committed golden
vectors from the reference's own kernels, the reference Triton kernels on the same GPU, and
size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): O, dQ, dK, dV within atol = rtol = 1e-2 of fp32-upcast SDPA;
LSE within 1e-3; error no worse than the reference Triton kernel's on the same inputs.
"""
import glob
import math
import os

import numpy as np
import pytest
import torch

import flashattn_b200 as fa
from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu
ATOL = RTOL = 1e-2
LSE_TOL = 1e-3
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _run_cuda(Q, K, V, dO, causal):
    q = Q.cuda().requires_grad_(True); k = K.cuda().requires_grad_(True); v = V.cuda().requires_grad_(True)
    O = fa.flash_attention(q, k, v, causal)            # the public entry (reference :169-170)
    O.backward(dO.cuda())
    _, LSE = fa.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal)
    torch.cuda.synchronize()
    return O.detach().cpu(), LSE.cpu(), q.grad.cpu(), k.grad.cpu(), v.grad.cpu()


def _close(a, b, atol=ATOL, rtol=RTOL):
    return torch.allclose(a.float(), b.float(), atol=atol, rtol=rtol)


def _norm_err(x, r, atol=ATOL, rtol=RTOL):
    """max |x - r| / (atol + rtol |r|): the contract holds iff this is <= 1 (code/_verify_func.py:29-31 uses the same quantity)."""
    r = r.float().to(x.device)
    return ((x.float() - r).abs() / (atol + rtol * r.abs())).max().item()


def _report(test, **kv):
    """Measured errors of the full-size / reference-comparison tests, appended to gpurun_out/parity_errors.jsonl (if writable)
    so that a GPU run leaves numbers behind, not just a verdict."""
    import json
    try:
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=test, **kv)) + "\n")
    except OSError:
        pass
    print(test, kv)


SMALL = [
    # B, H, Sq, Sk, D
    (1, 2, 128, 128, 64), (1, 2, 256, 256, 128), (2, 2, 384, 384, 64), (1, 3, 512, 512, 128),
    (1, 2, 128, 384, 64), (1, 2, 384, 128, 128),                      # cross attention, S_q != S_k
    (1, 2, 200, 300, 64), (1, 1, 333, 333, 128), (1, 2, 1, 77, 64),   # ragged / tiny (the reference mishandles these)
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["fp16", "bf16"])
@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("shape", SMALL, ids=lambda s: "x".join(map(str, s)))
def test_fwd_bwd_vs_oracle(shape, causal, dtype):
    B, H, Sq, Sk, D = shape
    Q, K, V, dO = orc.make_inputs(B, H, Sq, Sk, D, dtype, seed=Sq + Sk + D)
    O, LSE, dQ, dK, dV = _run_cuda(Q, K, V, dO, causal)
    rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, causal)          # fp64 ground truth
    sO, sdQ, sdK, sdV = orc.sdpa_fp32(Q, K, V, dO, causal)                   # the reference's yardstick
    assert (LSE - rLSE.float()).abs().max() < LSE_TOL
    for name, x, r, s in (("O", O, rO, sO), ("dQ", dQ, rdQ, sdQ), ("dK", dK, rdK, sdK), ("dV", dV, rdV, sdV)):
        assert torch.isfinite(x.float()).all(), name
        assert _close(x, r), f"{name} vs fp64 closed form: {(x.float() - r.float()).abs().max()}"
        assert _close(x, s), f"{name} vs fp32 SDPA"
    v = fa.verify_results(rO, O, rtol=1e-2, atol=1e-2)
    assert v["cosine_sim"] > 0.999                                            # code/_verify_func.py:37


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_against_reference_kernel_goldens(path):
    """Same seeded inputs as tests/golden/make_golden.py; outputs of the reference's Triton kernels."""
    z = np.load(path)
    m = {k[5:]: z[k].item() for k in z.files if k.startswith("meta_")}
    g = {k: torch.from_numpy(z[k]).float() for k in z.files if not k.startswith("meta_")}
    dt = getattr(torch, m["dtype"])
    Q, K, V, dO = orc.make_inputs(m["B"], m["H"], m["Sq"], m["Sk"], m["D"], dt, m["seed"])
    O, LSE, dQ, dK, dV = _run_cuda(Q, K, V, dO, bool(m["causal"]))
    assert (LSE - g["LSE"]).abs().max() < LSE_TOL
    # reference's own verdict rule (rtol=1e-2, atol=1e-3, cos>0.999) between the two 16-bit kernels; fp16 fixtures = the unmodified
    # kernels (CPU interpreter), bf16 fixtures = the bf16-patched kernels captured on a B200 (make_golden_gpu.py): 8x coarser ulp
    bound = 4e-3 if dt == torch.float16 else 3.2e-2
    for name, x in (("O", O), ("dQ", dQ), ("dK", dK), ("dV", dV)):
        r = fa.verify_results(g[name], x, name)
        assert r["cosine_sim"] > 0.9999 and r["max_abs_err"] < bound, (name, r)
    if "delta" in g:
        delta = fa.flash_attention_delta(O.cuda(), dO.cuda()).cpu()
        assert (delta - g["delta"]).abs().max() < 5e-3


def test_c1_config_vs_cpu_sdpa():
    """BASELINE config 1: B=1 H=4 N=512 D=64 non-causal, fp32 masters -> CPU SDPA oracle."""
    g = torch.Generator().manual_seed(0)
    Qm, Km, Vm, dOm = (torch.randn(1, 4, 512, 64, generator=g) for _ in range(4))
    sO, sdQ, sdK, sdV = orc.sdpa_fp32(Qm, Km, Vm, dOm, False)                 # fp32 master tensors
    for dt in (torch.float16, torch.bfloat16):
        Qd, Kd, Vd, dOd = Qm.to(dt), Km.to(dt), Vm.to(dt), dOm.to(dt)
        O, LSE, dQ, dK, dV = _run_cuda(Qd, Kd, Vd, dOd, False)
        # the contract (north star): within atol = rtol = 1e-2 of SDPA on the fp32-UPCAST of the 16-bit inputs the kernel saw
        uO, udQ, udK, udV = orc.sdpa_fp32(Qd, Kd, Vd, dOd, False)
        errs = {n: _norm_err(x, r) for n, x, r in (("O", O, uO), ("dQ", dQ, udQ), ("dK", dK, udK), ("dV", dV, udV))}
        # and against the fp32 MASTER tensors (SURVEY §8 C1 caveat): this adds the 16-bit rounding of the inputs themselves, which
        # is not the kernel's error — reported, bounded at twice the contract
        errs_master = {n: _norm_err(x, r) for n, x, r in (("O", O, sO), ("dQ", dQ, sdQ), ("dK", dK, sdK), ("dV", dV, sdV))}
        _report("C1", dtype=str(dt), norm_err_vs_upcast_inputs=errs, norm_err_vs_fp32_masters=errs_master)
        assert max(errs.values()) <= 1.0, errs
        assert max(errs_master.values()) <= 2.0, errs_master
        assert (LSE - orc.lse_bench(Qd, Kd, False)).abs().max() < LSE_TOL


def _gpu_truth(Q, K, V, dO, causal, q_chunk=4096):
    """fp32 attention forward + backward on the GPU with plain matmuls (TF32 off), chunked over query rows so that shapes far
    too big for the CPU oracle (C4: N = 8192, C5: N = 131072) fit: per (b, h) and q chunk S = q K^T is [q_chunk, S_k] fp32.
    Itself checked against the CPU oracle in test_gpu_truth_is_the_oracle.  Returns O, LSE, dQ, dK, dV (fp32)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, Sq, D = Q.shape
    Sk = K.shape[2]
    scale = 1 / math.sqrt(D)
    O = torch.empty(B, H, Sq, D, device=Q.device); LSE = torch.empty(B, H, Sq, device=Q.device)
    dQ = torch.empty_like(O); dK = torch.zeros(B, H, Sk, D, device=Q.device); dV = torch.zeros_like(dK)
    for b in range(B):
        for h in range(H):
            k, v = K[b, h].float(), V[b, h].float()
            for r0 in range(0, Sq, q_chunk):
                r1 = min(r0 + q_chunk, Sq)
                n = min(r1, Sk) if causal else Sk              # keys any row of the chunk can see
                q, do = Q[b, h, r0:r1].float(), dO[b, h, r0:r1].float()
                S = q @ k[:n].T * scale
                if causal:
                    i = torch.arange(r0, r1, device=Q.device); j = torch.arange(n, device=Q.device)
                    S.masked_fill_(~(i[:, None] >= j[None, :]), float("-inf"))
                lse = torch.logsumexp(S, -1); P = torch.exp(S - lse[:, None]); o = P @ v[:n]
                dP = do @ v[:n].T
                dS = P * (dP - (do * o).sum(-1, keepdim=True))
                O[b, h, r0:r1] = o; LSE[b, h, r0:r1] = lse
                dQ[b, h, r0:r1] = dS @ k[:n] * scale
                dK[b, h, :n] += dS.T @ q * scale; dV[b, h, :n] += P.T @ do
    return O, LSE, dQ, dK, dV


def test_gpu_truth_is_the_oracle():
    Q, K, V, dO = orc.make_inputs(1, 2, 256, 256, 64, torch.bfloat16, seed=1)
    t = _gpu_truth(Q.cuda(), K.cuda(), V.cuda(), dO.cuda(), True)
    r = orc.closed_form(Q, K, V, dO, True)
    for a, b in zip(t, r):
        assert (a.cpu() - b.float()).abs().max() < 1e-4


@pytest.mark.parametrize("cfg", [(4, 16, 2048, 64, True), (4, 16, 4096, 128, False), (2, 32, 8192, 128, True), (1, 1, 131072, 128, True)],
                         ids=["C2", "C3", "C4_shard", "C5_head"])
def test_baseline_configs_full_size(cfg):
    """BASELINE configs at full size, bf16, EVERY element against fp32 truth at the contract's atol = rtol = 1e-2 (the measured
    max normalised error is recorded): C2, C3, one 8-GPU shard of C4 (B = 16/8), and one head of C5 (N = 131072 — the (b,h)
    problems are independent, so a head is the whole difficulty of the long-context config).  Two (b,h) slices are additionally
    checked against the CPU oracle where that finishes in seconds."""
    B, H, S, D, causal = cfg
    g = torch.Generator(device="cuda").manual_seed(0)
    Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(4))
    q = Q.clone().requires_grad_(True); k = K.clone().requires_grad_(True); v = V.clone().requires_grad_(True)
    O = fa.flash_attention(q, k, v, causal); O.backward(dO)
    _, LSE = fa.flash_attention_forward(Q, K, V, causal)
    tO, tLSE, tdQ, tdK, tdV = _gpu_truth(Q, K, V, dO, causal)
    lse_err = (LSE - tLSE).abs().max().item()
    errs, frac_in = {}, {}
    for name, x, r in (("O", O, tO), ("dQ", q.grad, tdQ), ("dK", k.grad, tdK), ("dV", v.grad, tdV)):
        errs[name] = _norm_err(x.detach(), r)
        assert torch.nn.functional.cosine_similarity(x.detach().float().flatten(), r.flatten(), dim=0) > 0.9999, name
    _report("full_size", cfg=list(cfg), lse_abs_err=lse_err, max_norm_err=errs)
    assert lse_err < LSE_TOL
    assert max(errs.values()) <= 1.0, errs                  # atol = rtol = 1e-2 on every element, no tail allowance
    if S <= 4096:
        for (b, h) in ((0, 0), (B - 1, H - 1)):
            sl = lambda t: t[b:b + 1, h:h + 1].cpu()
            cO, cLSE = orc.sdpa_cpu_flash(sl(Q).float(), sl(K).float(), sl(V).float(), None, causal)
            assert _close(sl(O.detach()), cO) and (sl(LSE) - cLSE).abs().max() < LSE_TOL


def test_properties_at_full_size():
    """Size-independent properties on the C3 shape (B=4 H=16 N=4096 D=128 bf16)."""
    B, H, S, D = 4, 16, 4096, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    Q, K, V = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(3))
    # (1) rows of P sum to one: V = 1 -> O = 1
    O1, _ = fa.flash_attention_forward(Q, K, torch.ones_like(V), False)
    assert (O1.float() - 1).abs().max() < 8e-3
    # (2) determinism: bitwise equal across runs
    Oa, La = fa.flash_attention_forward(Q, K, V, True); Ob, Lb = fa.flash_attention_forward(Q, K, V, True)
    assert torch.equal(Oa, Ob) and torch.equal(La, Lb)
    # (3) causal prefix property: the first n rows do not depend on later keys — bitwise
    n = 1024
    Op, Lp = fa.flash_attention_forward(Q[:, :, :n].contiguous(), K[:, :, :n].contiguous(), V[:, :, :n].contiguous(), True)
    assert torch.equal(Op, Oa[:, :, :n]) and torch.equal(Lp, La[:, :, :n])
    # (4) (batch, head) independence: a head computed alone is bitwise the same (sharding invariant)
    Oh, Lh = fa.flash_attention_forward(Q[1:2, 3:4].contiguous(), K[1:2, 3:4].contiguous(), V[1:2, 3:4].contiguous(), True)
    assert torch.equal(Oh, Oa[1:2, 3:4]) and torch.equal(Lh, La[1:2, 3:4])
    # (5) key-permutation invariance (non-causal): O unchanged up to rounding, LSE to 1e-3
    perm = torch.randperm(S, device="cuda", generator=g)
    Of, Lf = fa.flash_attention_forward(Q, K, V, False)
    Oq, Lq = fa.flash_attention_forward(Q, K[:, :, perm].contiguous(), V[:, :, perm].contiguous(), False)
    assert (Lf - Lq).abs().max() < LSE_TOL and _close(Of, Oq)
    # (6) LSE merge over a key split == LSE over all keys (the ring invariant), via the CUDA merge kernel
    Oacc = torch.zeros(B, H, S, D, device="cuda"); Lacc = torch.full((B, H, S), float("-inf"), device="cuda")
    for lo in (0, S // 2):
        Opart, Lpart = fa.flash_attention_forward(Q, K[:, :, lo:lo + S // 2].contiguous(), V[:, :, lo:lo + S // 2].contiguous(), False)
        fa.merge_partial_(Oacc, Lacc, Opart, Lpart)
    assert (Lacc - Lf).abs().max() < LSE_TOL and _close(Oacc, Of)


def test_backward_is_deterministic_and_grads_flow():
    Q, K, V, dO = (t.cuda() for t in orc.make_inputs(2, 4, 1024, 1024, 128, torch.bfloat16, seed=4))
    O, LSE = fa.flash_attention_forward(Q, K, V, True)
    a = fa.flash_attention_backward(Q, K, V, O, dO, LSE, True)
    b = fa.flash_attention_backward(Q, K, V, O, dO, LSE, True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))           # two-kernel backward has no atomics
    # linearity of the backward in dO: scaling dO by 2 is exact in bf16 and delta = rowsum(dO o O) doubles exactly, so every
    # intermediate doubles exactly — the gradients must double BITWISE
    c = fa.flash_attention_backward(Q, K, V, O, (2 * dO.float()).bfloat16(), LSE, True)
    for x, y in zip(a, c):
        assert torch.equal((2 * x.float()).bfloat16(), y)


def test_sm_scale_and_noncontiguous_inputs():
    Q, K, V, dO = orc.make_inputs(1, 2, 256, 256, 64, torch.float16, seed=2)
    O = fa.flash_attention(Q.cuda(), K.cuda(), V.cuda(), True, sm_scale=0.2)
    rO, _ = orc.closed_form(Q, K, V, None, True, sm_scale=0.2)
    assert _close(O.cpu(), rO)
    Qb = Q.cuda().transpose(1, 2).contiguous().transpose(1, 2)    # [B,H,S,D] view of a [B,S,H,D] buffer
    assert not Qb.is_contiguous()
    O2 = fa.flash_attention(Qb, K.cuda(), V.cuda(), True, sm_scale=0.2)   # .contiguous() like reference :138
    assert torch.equal(O, O2)
    assert fa.attention is fa.flash_attention


def _norm_err_tensor(x, t, atol=1e-2, rtol=1e-2):
    return ((x.float() - t.float()).abs() / (atol + rtol * t.float().abs())).flatten()


def test_error_no_worse_than_reference_triton():
    """Same inputs through the reference Triton kernels (baseline/_ref: fp16 as shipped, bf16 with its dot-operand casts
    retargeted) and through this library; error against fp32 truth.  North star: "error no worse than the reference Triton
    kernel's".  Both implementations round P and dS to 16 bits at the same places, so the two error DISTRIBUTIONS coincide; the
    single largest of ~1e6 rounding errors is an order statistic of two equal distributions and which side holds it is a coin
    flip per tensor (measured: either way round, +-40 %, different tensors on different seeds).  Asserted per tensor, pooled over
    three seeds:
      rms error                       ours <= 1.05 x reference
      99.99th percentile (normalised) ours <= 1.10 x reference + 0.005
      maximum (normalised)            ours <= 1.0 (the contract, atol = rtol = 1e-2) and ours <= 2 x reference + 0.02 (no outliers)
    and every number, the maxima included, is recorded (gpurun_out/parity_errors.jsonl)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
    import ref_runner
    why = ref_runner.available()
    assert why is None, f"the reference copy (baseline/_ref) must travel with the repo: {why}"
    for dt, causal, D in ((torch.float16, True, 64), (torch.bfloat16, True, 64), (torch.bfloat16, False, 128), (torch.float16, True, 128)):
        ref_fn = ref_runner.ref_flash_attention(dt == torch.bfloat16)
        errs = {"ref": [[] for _ in range(4)], "ours": [[] for _ in range(4)]}
        sq = {"ref": [0.0] * 4, "ours": [0.0] * 4}
        for seed in (5, 6, 7):
            g = torch.Generator(device="cuda").manual_seed(seed)
            Q, K, V, dO = (torch.randn(2, 4, 1024, D, device="cuda", generator=g).to(dt) for _ in range(4))
            truth = _gpu_truth(Q, K, V, dO, causal)
            for name, fn in (("ref", ref_fn), ("ours", fa.flash_attention)):
                q = Q.clone().requires_grad_(True); k = K.clone().requires_grad_(True); v = V.clone().requires_grad_(True)
                O = fn(q, k, v, causal); O.backward(dO)
                for i, (x, t) in enumerate(zip((O.detach(), q.grad, k.grad, v.grad), (truth[0], truth[2], truth[3], truth[4]))):
                    errs[name][i].append(_norm_err_tensor(x, t))
                    sq[name][i] += (x.float() - t).pow(2).mean().item() / 3
        mx, q9999, rms = {}, {}, {}
        for name in ("ref", "ours"):
            pooled = [torch.cat(e) for e in errs[name]]
            mx[name] = [e.max().item() for e in pooled]
            q9999[name] = [e.kthvalue(int(0.9999 * e.numel())).values.item() for e in pooled]
            rms[name] = [v ** 0.5 for v in sq[name]]
        _report("vs_reference_triton", dtype=str(dt), causal=causal, D=D, tensors=["O", "dQ", "dK", "dV"], seeds=[5, 6, 7],
                max_norm_err=mx, q9999_norm_err=q9999, rms_err=rms)
        assert max(mx["ours"]) <= 1.0, mx                       # the contract itself
        for i in range(4):
            assert rms["ours"][i] <= 1.05 * rms["ref"][i] + 1e-7, (dt, causal, D, i, rms)
            assert q9999["ours"][i] <= 1.10 * q9999["ref"][i] + 0.005, (dt, causal, D, i, q9999)
            assert mx["ours"][i] <= 2.0 * mx["ref"][i] + 0.02, (dt, causal, D, i, mx)


def test_product_path_is_the_cuda_library():
    import flashattn_b200._cabi as cabi
    n0 = cabi.load().fa_sm100_launch_count()
    Q, K, V, dO = (t.cuda() for t in orc.make_inputs(1, 1, 128, 128, 64, torch.float16, seed=0))
    O, LSE = fa.flash_attention_forward(Q, K, V, False)
    fa.flash_attention_backward(Q, K, V, O, dO, LSE, False)
    assert cabi.load().fa_sm100_launch_count() - n0 == 4          # fwd, delta, fused dK/dV/dQ, dQ conversion (D = 64)
    assert cabi.last_hang() is None


def test_fused_backward_matches_two_kernel_backward():
    """Head dim 64: the fused single-pass backward (default) against the reference-structured two-kernel backward
    (deterministic mode) on the same inputs: same arithmetic up to the fused kernel's polynomial exp2 (rel. error 7.5e-5,
    a quarter of the elements) and dQ's fp32 summation order over kv tiles.  Both sit inside the oracle tolerance."""
    for (B, H, Hk, Sq, Sk, causal, dt) in ((2, 4, 4, 512, 512, True, torch.bfloat16), (1, 3, 3, 333, 333, True, torch.float16),
                                           (2, 8, 2, 320, 448, False, torch.bfloat16), (1, 2, 2, 200, 300, False, torch.float16)):
        g = torch.Generator().manual_seed(Sq + Sk)
        Q = torch.randn(B, H, Sq, 64, generator=g).to(dt); dO = torch.randn(B, H, Sq, 64, generator=g).to(dt)
        K = torch.randn(B, Hk, Sk, 64, generator=g).to(dt); V = torch.randn(B, Hk, Sk, 64, generator=g).to(dt)
        Qc, Kc, Vc, dOc = (t.cuda() for t in (Q, K, V, dO))
        O, LSE = fa.flash_attention_forward(Qc, Kc, Vc, causal)
        assert not fa.is_deterministic()
        fused = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, causal)
        prev = fa.set_deterministic(True)
        try:
            two = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, causal)
            two2 = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, causal)
        finally:
            fa.set_deterministic(prev)
        assert all(torch.equal(x, y) for x, y in zip(two, two2))              # deterministic mode is bitwise reproducible
        for x, y in zip(fused, two):           # a quarter of the fused kernel's exponentials come from the FMA-pipe polynomial:
            assert _close(x, y, 8e-3, 8e-3)    # results may differ by one 16-bit ulp (bf16: 2^-8 relative)
            assert (x.float() - y.float()).abs().mean() < 2e-4
        G = H // Hk
        _, _, rdQ, rdK, rdV = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, causal)
        assert _close(fused[0].cpu(), rdQ)
        if G == 1:
            assert _close(fused[1].cpu(), rdK) and _close(fused[2].cpu(), rdV)


def test_shared_sm_mode_is_bitwise_identical():
    """fa_sm100_set_shared_sms(1): persistent CTAs draw even their first item from the work counter (for launches that overlap
    NCCL transfers).  Which CTA computes which tile never changes a result: forward and the two-kernel backward are bitwise equal
    in both modes, at D = 64 and 128, also when the grid is larger than the number of items."""
    from flashattn_b200 import interface as I
    for (B, H, S, D, causal) in ((1, 2, 256, 64, True), (2, 8, 1024, 128, True), (4, 16, 2048, 64, False), (1, 1, 128, 128, False)):
        Q, K, V, dO = (t.cuda() for t in orc.make_inputs(B, H, S, S, D, torch.bfloat16, seed=B + S))
        prev_det = fa.set_deterministic(True)
        try:
            outs = []
            for mode in (False, True, False):
                prev = I.set_shared_sms(mode)
                try:
                    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
                    outs.append((O, LSE) + tuple(fa.flash_attention_backward(Q, K, V, O, dO, LSE, causal)))
                finally:
                    I.set_shared_sms(prev)
            for a, b in zip(outs[0], outs[1]):
                assert torch.equal(a, b)
            for a, b in zip(outs[0], outs[2]):      # and the counters were left clean for the next launch
                assert torch.equal(a, b)
        finally:
            fa.set_deterministic(prev_det)
    assert not I.set_shared_sms(False)


def test_first_call_on_a_thread_without_cuda_context():
    """cuTensorMapEncodeTiled is a driver call and needs a current context; a thread that has made no CUDA runtime call yet (the
    autograd worker running the first backward of a process, with every output coming from the caching allocator) has none.  The
    library binds the primary context itself (csrc/fa_api.cu::device_info): every entry point must work as the FIRST CUDA-related
    call of a fresh thread, with all tensors allocated beforehand.  (Found by the sweep: CUresult 201 on its first backward.)"""
    import threading
    from flashattn_b200 import interface as I
    for D in (64, 128):
        Q, K, V, dO = (t.cuda() for t in orc.make_inputs(1, 2, 256, 256, D, torch.bfloat16, seed=3))
        O, LSE = fa.flash_attention_forward(Q, K, V, True)
        ref = fa.flash_attention_backward(Q, K, V, O, dO, LSE, True)
        dQ, dK, dV = (torch.zeros_like(t) for t in (Q, K, V))
        delta = torch.empty(1, 2, 256, dtype=torch.float32, device="cuda"); acc = torch.empty(1, 2, 256, D, dtype=torch.float32, device="cuda")
        O2 = torch.empty_like(O); L2 = torch.empty_like(LSE)
        torch.cuda.synchronize()
        err = []

        def work():
            try:
                lib = __import__("flashattn_b200._cabi", fromlist=["load"]).load()
                st = torch.cuda.current_stream().cuda_stream
                rc = lib.fa_sm100_fwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O2.data_ptr(), L2.data_ptr(), 1, 2, 256, 256, D, 1, 1, 0.0, st)
                assert rc == 0, lib.fa_last_error()
                if D == 64:
                    I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, True, dq_acc=acc)
                else:
                    I.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, True, 7)
                torch.cuda.synchronize()
            except BaseException as e:   # noqa: BLE001
                err.append(e)

        for first in ("bwd", "fwd+bwd"):
            th = threading.Thread(target=work); th.start(); th.join()
            assert not err, err
        assert torch.equal(O2, O)
        for x, y in zip((dQ, dK, dV), ref):
            assert _close(x, y, 8e-3, 8e-3)


def test_fused128_backward_matches_two_kernel_backward_and_oracle():
    """Head dim 128: the opt-in fused single-pass backward (csrc/fa_bwd_fused128.cuh, FA_SM100_FUSED128=1 / set_fused128) against
    the default two-kernel backward on the same inputs.  Same exponentials and the same MMA order per kv tile, so dK and dV
    are BITWISE equal; dQ differs only by its fp32 summation order over kv tiles (TMA reduce-add) and the final rounding.
    Ragged shapes, S_q != S_k, GQA, both dtypes, a long causal sequence; every gradient also against the fp64 closed form."""
    from flashattn_b200 import interface as I
    for (B, H, Hk, Sq, Sk, causal, dt) in ((1, 1, 1, 128, 128, False, torch.bfloat16), (2, 4, 4, 512, 512, True, torch.bfloat16),
                                           (1, 3, 3, 333, 333, True, torch.float16), (2, 8, 2, 320, 448, False, torch.bfloat16),
                                           (1, 2, 2, 1, 77, False, torch.bfloat16), (1, 2, 2, 200, 300, False, torch.float16),
                                           (1, 2, 1, 2176, 2176, True, torch.bfloat16)):
        g = torch.Generator().manual_seed(Sq + Sk + H)
        Q = torch.randn(B, H, Sq, 128, generator=g).to(dt); dO = torch.randn(B, H, Sq, 128, generator=g).to(dt)
        K = torch.randn(B, Hk, Sk, 128, generator=g).to(dt); V = torch.randn(B, Hk, Sk, 128, generator=g).to(dt)
        Qc, Kc, Vc, dOc = (t.cuda() for t in (Q, K, V, dO))
        O, LSE = fa.flash_attention_forward(Qc, Kc, Vc, causal)
        two = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, causal)
        prev = I.set_fused128(True)
        try:
            assert I.fused_backward_supported(Qc)
            fused = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, causal)
        finally:
            I.set_fused128(prev)
        assert torch.equal(fused[1], two[1]) and torch.equal(fused[2], two[2]), (Sq, Sk)
        assert _close(fused[0], two[0], 8e-3, 8e-3) and (fused[0].float() - two[0].float()).abs().mean() < 2e-4
        G = H // Hk
        _, _, rdQ, rdK, rdV = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, causal)
        assert _close(fused[0].cpu(), rdQ), (Sq, Sk)
        if G == 1:
            assert _close(fused[1].cpu(), rdK) and _close(fused[2].cpu(), rdV)
    import flashattn_b200._cabi as cabi
    assert cabi.last_hang() is None


@pytest.mark.parametrize("shape", [(1, 3, 200, 300, 64), (2, 2, 333, 129, 128), (1, 2, 1, 77, 64), (1, 1, 640, 384, 128)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
def test_no_out_of_bounds_writes(shape, causal):
    """compute-sanitizer is closed on this pool, so: every output of the C-ABI calls sits between 4 KiB
    sentinel zones inside one arena; ragged shapes make every kernel run partial tiles (TMA clipping,
    predicated LSE/delta stores).  Any write outside a tensor's extent flips a sentinel."""
    import ctypes
    import flashattn_b200._cabi as cabi
    lib = cabi.load()
    B, H, Sq, Sk, D = shape
    GUARD = 4096
    sizes = dict(o=B * H * Sq * D * 2, lse=B * H * Sq * 4, dq=B * H * Sq * D * 2, dk=B * H * Sk * D * 2,
                 dv=B * H * Sk * D * 2, delta=B * H * Sq * 4, dq2=B * H * Sq * D * 2, dk2=B * H * Sk * D * 2,
                 dv2=B * H * Sk * D * 2, acc=B * H * Sq * D * 4)
    offs, cur = {}, GUARD
    for k, n in sizes.items():
        offs[k] = cur
        cur += (n + 255) // 256 * 256 + GUARD
    arena = torch.full((cur,), 0xA5, dtype=torch.uint8, device="cuda")
    Q, K, V, dO = (t.cuda() for t in orc.make_inputs(B, H, Sq, Sk, D, torch.bfloat16, seed=9))
    ptr = lambda k: arena.data_ptr() + offs[k]
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.fa_sm100_fwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), ptr("o"), ptr("lse"), B, H, Sq, Sk, D, 1, int(causal), 0.0, st)
    assert rc == 0, lib.fa_last_error()
    rc = lib.fa_sm100_bwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), ptr("o"), dO.data_ptr(), ptr("lse"), ptr("dq"), ptr("dk"),
                          ptr("dv"), ptr("delta"), B, H, Sq, Sk, D, 1, int(causal), 0.0, st)
    assert rc == 0, lib.fa_last_error()
    if D == 64:           # the fused backward (TMA reduce-add into the fp32 workspace, conversion kernel) into its own outputs
        assert lib.fa_sm100_bwd_fused_workspace(B, H, Sq, D) == sizes["acc"]
        rc = lib.fa_sm100_bwd_fused(Q.data_ptr(), K.data_ptr(), V.data_ptr(), ptr("o"), dO.data_ptr(), ptr("lse"), ptr("dq2"),
                                    ptr("dk2"), ptr("dv2"), ptr("delta"), ptr("acc"), B, H, H, Sq, Sk, D, 1, int(causal), 0.0,
                                    None, st, 0)
        assert rc == 0, lib.fa_last_error()
    torch.cuda.synchronize()
    mask = torch.ones(cur, dtype=torch.bool, device="cuda")
    for k, n in sizes.items():
        mask[offs[k]:offs[k] + n] = False
    assert bool((arena[mask] == 0xA5).all()), "a kernel wrote outside its output tensor"
    # and the in-bounds results are the right ones
    O = arena[offs["o"]:offs["o"] + sizes["o"]].view(torch.bfloat16).view(B, H, Sq, D)
    dK = arena[offs["dk"]:offs["dk"] + sizes["dk"]].view(torch.bfloat16).view(B, H, Sk, D)
    rO, _, _, rdK, _ = orc.closed_form(Q.cpu(), K.cpu(), V.cpu(), dO.cpu(), causal)
    assert _close(O.cpu(), rO) and _close(dK.cpu(), rdK)
    if D == 64:
        dQ2 = arena[offs["dq2"]:offs["dq2"] + sizes["dq2"]].view(torch.bfloat16).view(B, H, Sq, D)
        dK2 = arena[offs["dk2"]:offs["dk2"] + sizes["dk2"]].view(torch.bfloat16).view(B, H, Sk, D)
        rdQ = orc.closed_form(Q.cpu(), K.cpu(), V.cpu(), dO.cpu(), causal)[2]
        assert _close(dQ2.cpu(), rdQ) and _close(dK2, dK, 8e-3, 8e-3)      # one bf16 ulp
    assert cabi.last_hang() is None


@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("D", [64, 128])
def test_bshd_strided_inputs_zero_copy(D, causal):
    """[B,S,H,D]-layout buffers (what a fused QKV projection produces) go through strided tensor maps: same
    numbers as the contiguous path (bitwise), no .contiguous() copy, outputs and grads in the caller's layout."""
    B, S, H = 2, 384, 3
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B, S, 3, H, D, device="cuda", generator=g).bfloat16()        # packed projection output
    q, k, v = (qkv[:, :, i].requires_grad_(True) for i in range(3))                # [B,S,H,D] views, row stride 3*H*D
    dO = torch.randn(B, S, H, D, device="cuda", generator=g).bfloat16()
    assert fa.tma_compatible(q.transpose(1, 2)) and not q.transpose(1, 2).is_contiguous()
    O = fa.flash_attention_bshd(q, k, v, causal)
    assert O.shape == (B, S, H, D)
    O.backward(dO)
    qc, kc, vc = (t.detach().transpose(1, 2).contiguous().requires_grad_(True) for t in (q, k, v))
    Oc = fa.flash_attention(qc, kc, vc, causal)
    Oc.backward(dO.transpose(1, 2).contiguous())
    assert torch.equal(O.transpose(1, 2), Oc)
    for name, a, b in (("dQ", q.grad, qc.grad), ("dK", k.grad, kc.grad), ("dV", v.grad, vc.grad)):
        if name == "dQ" and D == 64:          # fused backward: dQ's fp32 summation order over kv tiles is not fixed
            assert _close(a.transpose(1, 2), b, 8e-3, 8e-3)
        else:
            assert torch.equal(a.transpose(1, 2), b)
    rO, _, rdQ, _, _ = orc.closed_form(qc.detach().cpu(), kc.detach().cpu(), vc.detach().cpu(), dO.transpose(1, 2).cpu(), causal)
    assert _close(Oc.detach().cpu(), rO) and _close(qc.grad.cpu(), rdQ)
    # a layout the TMA cannot express (row stride not a multiple of 8 elements) falls back to the reference's copy
    odd = torch.randn(1, 2, 128, D + 4, device="cuda").bfloat16()[..., :D]
    assert not fa.tma_compatible(odd)
    assert _close(fa.flash_attention(odd, odd, odd, causal).cpu(),
                  orc.closed_form(odd.cpu(), odd.cpu(), odd.cpu(), None, causal)[0])


@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("H,Hk,D", [(8, 2, 64), (4, 1, 128), (6, 3, 128)], ids=["gqa4", "mqa", "gqa2"])
def test_gqa_mqa_head_sharing(H, Hk, D, causal):
    """K/V with Hk heads shared by groups of H/Hk query heads: equals attention with K/V repeated per query head,
    and dK/dV equal the group sums (reduced inside the dK/dV kernel, deterministic)."""
    B, Sq, Sk = 2, 320, 448 if not causal else 320
    g = torch.Generator().manual_seed(21)
    Q = torch.randn(B, H, Sq, D, generator=g).bfloat16(); dO = torch.randn(B, H, Sq, D, generator=g).bfloat16()
    K = torch.randn(B, Hk, Sk, D, generator=g).bfloat16(); V = torch.randn(B, Hk, Sk, D, generator=g).bfloat16()
    q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    O = fa.flash_attention(q, k, v, causal); O.backward(dO.cuda())
    assert k.grad.shape == K.shape and v.grad.shape == V.shape
    G = H // Hk
    Ke, Ve = K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1)
    rO, _, rdQ, rdKe, rdVe = orc.closed_form(Q, Ke, Ve, dO, causal)
    rdK = rdKe.reshape(B, Hk, G, Sk, D).sum(2); rdV = rdVe.reshape(B, Hk, G, Sk, D).sum(2)
    for name, x, r in (("O", O, rO), ("dQ", q.grad, rdQ), ("dK", k.grad, rdK), ("dV", v.grad, rdV)):
        # dK/dV are sums over the G query heads of a group, each term within the contract: the sum of G independent rounding
        # errors is held to atol = 1e-2 * sqrt(G) (rtol stays 1e-2); O and dQ to the contract itself
        atol = 1e-2 if name in ("O", "dQ") else 1e-2 * math.sqrt(G)
        assert _close(x.detach().cpu(), r, atol, 1e-2), (name, (x.detach().cpu().float() - r.float()).abs().max())
    # bitwise the same as running the expanded problem's forward through the non-GQA path
    Oe = fa.flash_attention(Q.cuda(), Ke.cuda(), Ve.cuda(), causal)
    assert torch.equal(O.detach(), Oe)
    k2, v2 = (t.cuda().requires_grad_(True) for t in (K, V))
    q2 = Q.cuda().requires_grad_(True)
    fa.flash_attention(q2, k2, v2, causal).backward(dO.cuda())
    assert torch.equal(k2.grad, k.grad) and torch.equal(v2.grad, v.grad)      # deterministic


def _check_ranges(Q, K, V, dO, causal, ranges, tol_kv=None):
    """CUDA path with a Ranges mask against the fp64 closed form with the same mask (dense, on the CPU)."""
    q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    O = fa.flash_attention(q, k, v, causal, ranges=ranges)
    O.backward(dO.cuda())
    _, LSE = fa.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal, ranges=ranges)
    G = Q.shape[1] // K.shape[1]
    rO, rLSE, rdQ, rdKe, rdVe = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, causal,
                                                row_ranges=(ranges.row_lo.cpu(), ranges.row_hi.cpu()))
    B, Hk, Sk, D = K.shape
    rdK = rdKe.reshape(B, Hk, G, Sk, D).sum(2); rdV = rdVe.reshape(B, Hk, G, Sk, D).sum(2)
    assert (LSE.cpu() - rLSE.float()).abs().max() < LSE_TOL
    # dK/dV of a K/V head shared by G query heads are G-term sums: atol = 1e-2 * sqrt(G) (see test_gqa_mqa_head_sharing)
    tol_kv = 1e-2 * math.sqrt(G) if tol_kv is None else tol_kv
    for name, x, r, tol in (("O", O, rO, 1e-2), ("dQ", q.grad, rdQ, 1e-2), ("dK", k.grad, rdK, tol_kv), ("dV", v.grad, rdV, tol_kv)):
        x = x.detach().cpu()
        assert torch.isfinite(x.float()).all(), name
        assert _close(x, r, tol, 1e-2), (name, (x.float() - r.float()).abs().max().item())
    import flashattn_b200._cabi as cabi
    assert cabi.last_hang() is None
    return O.detach()


@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("D,H,Hk", [(64, 4, 4), (128, 4, 2)], ids=["d64", "d128gqa"])
def test_varlen_packed_sequences(D, H, Hk, causal):
    """Packed variable-length self-attention (cu_seqlens; Phase_6.md:160-174 lists it as future work): block-diagonal mask
    through Ranges, zero-copy [total, H, D] layout, tiles outside a sequence skipped.  Sequence boundaries fall inside tiles,
    on tile edges and inside one 256-row forward item."""
    lens = [100, 333, 27, 512, 128, 1, 250]
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    total = cu[-1]
    g = torch.Generator().manual_seed(31)
    q = torch.randn(total, H, D, generator=g).bfloat16(); do = torch.randn(total, H, D, generator=g).bfloat16()
    k = torch.randn(total, Hk, D, generator=g).bfloat16(); v = torch.randn(total, Hk, D, generator=g).bfloat16()
    ranges = fa.Ranges.from_cu_seqlens(cu, total, device="cuda")
    to4 = lambda t: t.transpose(0, 1)[None]
    O = _check_ranges(to4(q), to4(k), to4(v), to4(do), causal, ranges)
    # the packed entry point returns the same thing in the packed layout
    Op = fa.flash_attention_varlen(q.cuda(), k.cuda(), v.cuda(), torch.tensor(cu), causal)
    assert Op.shape == (total, H, D) and torch.equal(Op, O[0].transpose(0, 1))
    # and every sequence equals the plain operator run on that sequence alone (up to tile-alignment rounding)
    for s in (1, 3):
        sl = slice(cu[s], cu[s + 1])
        Os = fa.flash_attention(to4(q[sl]).cuda().contiguous(), to4(k[sl]).cuda().contiguous(), to4(v[sl]).cuda().contiguous(), causal)
        assert _close(Os[0].transpose(0, 1), Op[sl], 4e-3, 4e-3)


@pytest.mark.parametrize("D", [64, 128])
def test_key_padding_and_sliding_window(D):
    g = torch.Generator().manual_seed(33)
    B, H, S = 2, 3, 640
    Q, K, V, dO = (torch.randn(B, H, S, D, generator=g).bfloat16() for _ in range(4))
    # key padding: batch 0 has 200 valid keys, batch 1 all of them; padded keys get exactly zero gradient
    r = fa.Ranges.from_key_padding([200, S], S, S, device="cuda")
    _check_ranges(Q, K, V, dO, False, r)
    k = K.cuda().requires_grad_(True); v = V.cuda().requires_grad_(True)
    fa.flash_attention(Q.cuda(), k, v, False, ranges=r).backward(dO.cuda())
    assert (k.grad[0, :, 200:] == 0).all() and (v.grad[0, :, 200:] == 0).all()
    # causal sliding window of 200 keys: only ~1/3 of the causal tiles are visited
    _check_ranges(Q, K, V, dO, True, fa.Ranges.sliding_window(B, S, 200, device="cuda"))


@pytest.mark.parametrize("D", [64, 128])
def test_empty_items_do_not_race_with_the_previous_store(D):
    """Work items with nothing to do (kv tiles no query sees, q tiles whose rows see no key) reach their epilogue at once.  At
    D = 128 the output staging aliases the resident operand tiles, and until round 2 only thread 0 waited for the PREVIOUS item's
    TMA store to have read it: the other threads zeroed the staging under that store, so the tile in front of an empty item
    came out partly zero — different from run to run (seen as a drifting dK checksum in the sequence-parallel stress run).  Many
    empty items behind few full ones, twenty runs: every run bitwise equal to the first and inside the contract."""
    g = torch.Generator().manual_seed(91)
    B, H, Sq, Sk = 1, 8, 2048, 8192
    Q = torch.randn(B, H, Sq, D, generator=g).bfloat16(); dO = torch.randn(B, H, Sq, D, generator=g).bfloat16()
    K = torch.randn(B, H, Sk, D, generator=g).bfloat16(); V = torch.randn(B, H, Sk, D, generator=g).bfloat16()
    # keys [0, 1000) visible to query rows [0, 1500); the other 56 kv tiles and the last 4 q tiles are empty items
    i = torch.arange(Sq); j = torch.arange(Sk)
    row_lo = torch.where(i < 1500, 0, 1000)[None].expand(B, Sq); row_hi = torch.full((B, Sq), 1000, dtype=torch.int64)   # monotone
    col_lo = torch.where(j < 1000, 0, 1500)[None].expand(B, Sk); col_hi = torch.full((B, Sk), 1500, dtype=torch.int64)
    r = fa.Ranges(row_lo.cuda(), row_hi.cuda(), col_lo.cuda(), col_hi.cuda()).validate()
    Qc, Kc, Vc, dOc = (t.cuda() for t in (Q, K, V, dO))
    prev = fa.set_deterministic(True)
    try:
        O, LSE = fa.flash_attention_forward(Qc, Kc, Vc, False, None, r)
        first = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, False, None, r)
        for _ in range(20):
            again = fa.flash_attention_backward(Qc, Kc, Vc, O, dOc, LSE, False, None, r)
            for a, b in zip(first, again):
                assert torch.equal(a, b)
    finally:
        fa.set_deterministic(prev)
    rO, _, rdQ, rdK, rdV = orc.closed_form(Q[:, :, :1500], K[:, :, :1000], V[:, :, :1000], dO[:, :, :1500], False)
    assert _close(O[:, :, :1500].cpu(), rO) and (O[:, :, 1500:] == 0).all()
    assert _close(first[0][:, :, :1500].cpu(), rdQ) and (first[0][:, :, 1500:] == 0).all()
    assert _close(first[1][:, :, :1000].cpu(), rdK) and (first[1][:, :, 1000:] == 0).all()
    assert _close(first[2][:, :, :1000].cpu(), rdV) and (first[2][:, :, 1000:] == 0).all()


@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("D,H,Hk,p", [(64, 4, 4, 0.25), (128, 4, 2, 0.1)], ids=["d64", "d128gqa"])
def test_dropout_matches_oracle_mask(D, H, Hk, p, causal):
    """Dropout on the attention probabilities (Phase_6.md:54-114 lists it as future work).  The keep mask is a pure function of
    (seed, batch*head, query, key); the oracle restates it in numpy, so O, dQ, dK, dV are compared element for element against
    the fp64 closed form with the SAME mask; the forward and both backward kernels regenerate it in different orientations."""
    B, Sq, Sk, seed = 2, 320, 320 if causal else 448, 0x1234_5678_9ABC_DEF0
    g = torch.Generator().manual_seed(41)
    Q = torch.randn(B, H, Sq, D, generator=g).bfloat16(); dO = torch.randn(B, H, Sq, D, generator=g).bfloat16()
    K = torch.randn(B, Hk, Sk, D, generator=g).bfloat16(); V = torch.randn(B, Hk, Sk, D, generator=g).bfloat16()
    q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    O = fa.flash_attention(q, k, v, causal, dropout_p=p, dropout_seed=seed)
    O.backward(dO.cuda())
    keep, scale = orc.dropout_keep_mask(seed, B, H, Sq, Sk, p)
    assert abs(keep.float().mean().item() - (1 - round(p * 256) / 256)) < 5e-3
    G = H // Hk
    rO, rLSE, rdQ, rdKe, rdVe = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, causal,
                                                keep_mask=keep, keep_scale=scale)
    rdK = rdKe.reshape(B, Hk, G, Sk, D).sum(2); rdV = rdVe.reshape(B, Hk, G, Sk, D).sum(2)
    for name, x, r in (("O", O, rO), ("dQ", q.grad, rdQ), ("dK", k.grad, rdK), ("dV", v.grad, rdV)):
        # gradients: 2.5e-2.  delta = rowsum(dO o O) uses the 16-bit O like the reference (kernel :210-211); with the 1/(1-p) factor
        # O is no longer exactly representable where a row sees a single key, so dS = P (dP - delta) keeps a rounding residue
        # (measured 2.1e-2 on one element of causal row 0) where exact arithmetic cancels to zero.
        tol = (1e-2 if G == 1 else 2e-2) if name == "O" else 2.5e-2
        assert _close(x.detach().cpu(), r, tol, tol), (name, (x.detach().cpu().float() - r.float()).abs().max().item())
    # LSE is that of the undropped softmax; same seed -> bitwise the same; another seed -> another mask; p = 0 -> the plain operator
    _, L0 = fa.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal)
    O1, L1 = fa.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal, dropout_p=p, dropout_seed=seed)
    O2, _ = fa.flash_attention_forward(q.detach(), k.detach(), v.detach(), causal, dropout_p=p, dropout_seed=seed + 1)
    assert torch.equal(L0, L1) and torch.equal(O1, O.detach()) and not torch.equal(O1, O2)
    assert torch.equal(fa.flash_attention(q.detach(), k.detach(), v.detach(), causal, dropout_p=0.0),
                       fa.flash_attention(q.detach(), k.detach(), v.detach(), causal))
    # dropout + packed variable-length sequences together (B = 1 view of the first batch entry)
    if D == 64:
        r = fa.Ranges.from_cu_seqlens([0, 100, 320], Sq, device="cuda")
        Or = fa.flash_attention(q.detach()[:1], k.detach()[:1, :, :Sq], v.detach()[:1, :, :Sq], causal, ranges=r, dropout_p=p, dropout_seed=seed)
        rr, _ = orc.closed_form(Q[:1], K[:1, :, :Sq], V[:1, :, :Sq], None, causal, row_ranges=(r.row_lo.cpu(), r.row_hi.cpu()),
                                keep_mask=keep[:1, :, :, :Sq], keep_scale=scale)
        assert _close(Or.cpu(), rr)


def test_native_host_path_equals_python_path():
    """flash_attention() runs the C++ autograd node (csrc/fa_torch.cpp, _fa_torch.so) for the plain operator; the
    reference-shaped Python class FlashAttentionFunction is the same operator.  Same kernels, same arguments: bitwise equal
    (head dim 64: in deterministic mode, the fused backward's dQ summation order is scheduling dependent)."""
    import flashattn_b200.interface as itf
    import flashattn_b200._cabi as cabi
    assert itf._HOST_NATIVE and itf._load_native() is not None
    for D, dt, causal in ((128, torch.bfloat16, True), (64, torch.float16, False), (64, torch.bfloat16, True)):
        Q, K, V, dO = (t.cuda() for t in orc.make_inputs(2, 4, 384, 384, D, dt, seed=D))
        prev = fa.set_deterministic(True)
        try:
            res = []
            n0 = cabi.load().fa_sm100_launch_count()
            for fn in (lambda q, k, v: fa.flash_attention(q, k, v, causal),
                       lambda q, k, v: fa.FlashAttentionFunction.apply(q, k, v, causal)):
                q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
                O = fn(q, k, v); O.backward(dO)
                res.append((O.detach(), q.grad, k.grad, v.grad))
            assert cabi.load().fa_sm100_launch_count() - n0 == 8      # (fwd, delta, dQ, dK/dV) x 2: both paths launch the library
        finally:
            fa.set_deterministic(prev)
        for a, b in zip(*res):
            assert torch.equal(a, b)
    # sm_scale, GQA and a non-contiguous input (falls back to the strided Python path, same result)
    Q, K, V, dO = (t.cuda() for t in orc.make_inputs(1, 4, 256, 256, 128, torch.bfloat16, seed=3))
    O1 = fa.flash_attention(Q, K[:, :2].contiguous(), V[:, :2].contiguous(), True, sm_scale=0.1)
    O2 = fa.FlashAttentionFunction.apply(Q, K[:, :2].contiguous(), V[:, :2].contiguous(), True, 0.1)
    assert torch.equal(O1, O2)
    with pytest.raises(AssertionError):                                   # the reference's asserts (:133-136) still fire first
        fa.flash_attention(Q.float(), K.float(), V.float())


def test_cuda_graph_capture_replays_bitwise():
    """fwd + bwd captured in a CUDA graph and replayed: no host work per step.  Possible because nothing on the path syncs or
    memsets (the persistent kernels' work counters reset themselves) and outputs come from the capture's private pool.
    Replays must reproduce the eager results bitwise (D = 128: every tensor; D = 64: O, dK, dV — the fused backward's dQ is
    summed in scheduling order) and follow in-place changes of the inputs."""
    for D in (128, 64):
        Q, K, V, dO = (t.cuda() for t in orc.make_inputs(2, 4, 512, 512, D, torch.bfloat16, seed=11 + D))
        q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))

        def eager():
            O = fa.flash_attention(q, k, v, True)
            return (O.detach().clone(),) + tuple(x.clone() for x in torch.autograd.grad(O, (q, k, v), dO))
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                     # warm-up off the capture (kernel attributes, allocator)
            for _ in range(3):
                eager()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            O = fa.flash_attention(q, k, v, True)
            gq, gk, gv = torch.autograd.grad(O, (q, k, v), dO)
        for trial in range(3):
            if trial == 2:                                # new data in the captured input buffers
                with torch.no_grad():
                    q.copy_(torch.roll(Q, 1, 2)); k.mul_(0.5)
            ref = eager()
            graph.replay(); torch.cuda.synchronize()
            assert torch.equal(O.detach(), ref[0]) and torch.equal(gk, ref[2]) and torch.equal(gv, ref[3]), (D, trial)
            if D == 128:
                assert torch.equal(gq, ref[1]), trial
            else:
                assert _close(gq, ref[1])
        del graph


@pytest.mark.parametrize("D", [64, 128])
def test_varlen_padded_tail_is_never_read(D):
    """A packed buffer longer than cu_seqlens[-1]: the kernels get views of the packed tokens only, so whatever the tail holds
    (NaN here) never reaches a real token's output or gradient, pad outputs / gradients are zero, and the real part is bitwise
    the unpadded call."""
    H, cu, total = 2, [0, 200, 333], 448
    g = torch.Generator().manual_seed(5)
    q, k, v, do = (torch.randn(total, H, D, generator=g).bfloat16().cuda() for _ in range(4))
    for t in (q, k, v, do):
        t[cu[-1]:] = float("nan")
    r = fa.Ranges.from_cu_seqlens(cu, total, device="cuda").validate()
    assert r.n_tokens == cu[-1] and r.n_buffer == total
    q2, k2, v2 = (t.clone().requires_grad_(True) for t in (q, k, v))
    O = fa.flash_attention_varlen(q2, k2, v2, torch.tensor(cu), False)
    assert O.shape == q.shape
    O.backward(torch.nan_to_num(do))
    q3, k3, v3 = (t[:cu[-1]].clone().requires_grad_(True) for t in (q, k, v))
    O3 = fa.flash_attention_varlen(q3, k3, v3, torch.tensor(cu), False, ranges=r)
    O3.backward(do[:cu[-1]])
    for a, b in ((O, O3), (q2.grad, q3.grad), (k2.grad, k3.grad), (v2.grad, v3.grad)):
        assert torch.isfinite(a.float()).all()
        assert torch.equal(a[:cu[-1]].detach(), b.detach()) and (a[cu[-1]:] == 0).all()


@pytest.mark.parametrize("D", [64, 128])
@pytest.mark.parametrize("n_empty", [128, 256, 300])
def test_rows_with_empty_key_range(D, n_empty):
    """Hand-built Ranges in which the first n_empty query rows see NO key (a whole 128-row tile, a whole 256-row forward item,
    and a ragged count): those rows get O = 0, LSE = -inf and dQ = 0 — defined values, not the contents of torch.empty — and
    every other row equals attention over the visible part."""
    B, H, S = 1, 2, 640
    Q, K, V, dO = (t.cuda() for t in orc.make_inputs(B, H, S, S, D, torch.bfloat16, seed=n_empty + D))
    i = torch.arange(S)
    row_lo = torch.zeros(B, S, dtype=torch.int32); row_hi = torch.where(i < n_empty, 0, S).to(torch.int32)[None]
    col_lo = torch.full((B, S), n_empty, dtype=torch.int32); col_hi = torch.full((B, S), S, dtype=torch.int32)
    r = fa.Ranges(row_lo.cuda(), row_hi.cuda(), col_lo.cuda(), col_hi.cuda())
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    O = fa.flash_attention(q, k, v, False, ranges=r); O.backward(dO)
    _, LSE = fa.flash_attention_forward(Q, K, V, False, ranges=r)
    assert (O[:, :, :n_empty] == 0).all() and (LSE[:, :, :n_empty] == float("-inf")).all() and (q.grad[:, :, :n_empty] == 0).all()
    rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q[:, :, n_empty:].cpu(), K.cpu(), V.cpu(), dO[:, :, n_empty:].cpu(), False)
    assert (LSE[:, :, n_empty:].cpu() - rLSE.float()).abs().max() < LSE_TOL
    for name, x, ref in (("O", O[:, :, n_empty:], rO), ("dQ", q.grad[:, :, n_empty:], rdQ), ("dK", k.grad, rdK), ("dV", v.grad, rdV)):
        assert _close(x.detach().cpu(), ref), name
    import flashattn_b200._cabi as cabi
    assert cabi.last_hang() is None


def test_reference_self_check_compare_with_sdpa():
    """The reference's own parity routine (code/My_FlashAttention_optimized.py:172-212), run against this library through the
    drop-in module: SDPA flash backend as yardstick, verify_results verdicts."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "flashattention-from-scratch-with-triton_b200",
                        "My_FlashAttention_optimized.py")
    spec = importlib.util.spec_from_file_location("My_FlashAttention_optimized_dropin", path)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    g = torch.Generator(device="cuda").manual_seed(42)
    Q, K, V = (torch.randn(4, 8, 256, 64, device="cuda", generator=g, dtype=torch.float16) for _ in range(3))   # the __main__ shape
    res = mod.compare_with_sdpa(Q, K, V, True, verbose=False)
    assert [r["name"] for r in res] == ["O", "dQ", "dK", "dV"] and all(r["passed"] for r in res), res
