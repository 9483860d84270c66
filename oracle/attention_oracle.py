"""CPU oracle for the flash-attention hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never does: it
calls the sm_100a CUDA library through the C ABI and fails loudly if that is missing.

What is restated here (all citations relative to /root/reference):

* ``forward_blocked``  – code/_flash_attention_kernel_optimized.py:60-129  
* ``dq_blocked``       – code/_flash_attention_kernel_optimized.py:188-258 (App. A.2)
* ``dkv_blocked``      – code/_flash_attention_kernel_optimized.py:315-386 (App. A.3)
* ``closed_form``      – Phase_4.md:1251-1271 (gradient identities), Phase_3.md:699-708 (LSE)
* ``sdpa_fp32``        – the reference's own ground truth: fp32-upcast
                         ``F.scaled_dot_product_attention`` (Phase_3.md:244-252,
                         code/My_FlashAttention_optimized.py:178-187)

Pinning: the reference ships no golden vectors.  The blocked restatement is
pinned against outputs of the *reference's own Triton kernels* executed by the Triton CPU
interpreter in the build container (``tests/golden/make_golden.py`` → ``tests/golden/*.npz``);
``tests/test_oracle.py`` checks oracle == golden.  bf16 goldens come from a copy of the
reference with its four hard ``tl.float16`` dot-operand casts retargeted to bf16 (the
shipped reference asserts on bf16), and are labelled "patched" in the fixture metadata.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

LOG2_E = 1.44269504  # code/_flash_attention_kernel_optimized.py:79 (fp32 literal in the kernel)


def _r16(x: torch.Tensor, dt: torch.dtype) -> torch.Tensor:
    """Round an fp32 tensor to the 16-bit run dtype and come back to fp32."""
    return x.to(dt).to(torch.float32)


def _exp2(x: torch.Tensor) -> torch.Tensor:
    return torch.exp2(x)


def forward_blocked(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, is_causal: bool,
                    BLOCK_M: int = 64, BLOCK_N: int = 64,
                    sm_scale: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Online-softmax forward, tile by tile, exactly as the reference kernel orders it.

    Follows code/_flash_attention_kernel_optimized.py:60-129.  Q,K,V: [B,H,S,D] fp16/bf16 on
    CPU.  Returns O (input dtype) and LSE (fp32, natural log of the scaled scores).
    """
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    dt = Q.dtype
    scale = torch.tensor(1.0 / (D ** 0.5) if sm_scale is None else sm_scale, dtype=torch.float32)
    c = torch.tensor(LOG2_E, dtype=torch.float32)
    Qf, Kf, Vf = Q.float(), K.float(), V.float()
    O = torch.zeros(B, H, S_q, D, dtype=dt)
    LSE = torch.zeros(B, H, S_q, dtype=torch.float32)
    neg_inf = float("-inf")
    for q0 in range(0, S_q, BLOCK_M):
        q1 = min(q0 + BLOCK_M, S_q)
        rows = torch.arange(q0, q1)
        Qb = Qf[:, :, q0:q1]                                   # :71-72
        m = torch.full((B, H, q1 - q0), neg_inf)               # :75
        l = torch.zeros(B, H, q1 - q0)                         # :76
        o = torch.zeros(B, H, q1 - q0, D)                      # :77
        loop_end = min(q0 + BLOCK_M, S_k) if is_causal else S_k  # :82 (loads past S_k are all masked)
        for s0 in range(0, loop_end, BLOCK_N):
            s1 = min(s0 + BLOCK_N, S_k)
            cols = torch.arange(s0, s1)
            S = torch.matmul(Qb, Kf[:, :, s0:s1].transpose(-1, -2)) * scale     # :93
            if is_causal and not (q0 >= s0 + BLOCK_N - 1):                      # :98-101
                S = torch.where(rows[:, None] >= cols[None, :], S, torch.tensor(neg_inf))  # :102-103
            m_new = torch.maximum(m, S.max(dim=-1).values)                      # :106
            corr = _exp2((m - m_new) * c)                                       # :108
            p = _exp2((S - m_new[..., None]) * c)                               # :109
            l = l * corr + p.sum(dim=-1)                                        # :111 (fp32 p)
            o = o * corr[..., None] + torch.matmul(_r16(p, dt), Vf[:, :, s0:s1])  # :115
            m = m_new
        O[:, :, q0:q1] = (o / l[..., None]).to(dt)                              # :120-123
        LSE[:, :, q0:q1] = m + torch.log(l)                                     # :126-129
    return O, LSE


def _p_block(Qb, Kb, lse_b, rows, cols, scale, c, is_causal, diag):
    S = torch.matmul(Qb, Kb.transpose(-1, -2)) * scale
    if is_causal and diag:
        S = torch.where(rows[:, None] >= cols[None, :], S, torch.tensor(float("-inf")))
    return _exp2((S - lse_b[..., None]) * c)


def dq_blocked(Q, K, V, O, dO, LSE, is_causal: bool, BLOCK_M: int = 64, BLOCK_N: int = 64,
               sm_scale: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """dQ and delta, following code/_flash_attention_kernel_optimized.py:188-258."""
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    dt = Q.dtype
    scale = torch.tensor(1.0 / (D ** 0.5) if sm_scale is None else sm_scale, dtype=torch.float32)
    c = torch.tensor(LOG2_E, dtype=torch.float32)
    Qf, Kf, Vf, Of, dOf = Q.float(), K.float(), V.float(), O.float(), dO.float()
    dQ = torch.zeros(B, H, S_q, D, dtype=dt)
    delta = (dOf * Of).sum(dim=-1)                                              # :210-211
    for q0 in range(0, S_q, BLOCK_M):
        q1 = min(q0 + BLOCK_M, S_q)
        rows = torch.arange(q0, q1)
        acc = torch.zeros(B, H, q1 - q0, D)
        loop_end = min(q0 + BLOCK_M, S_k) if is_causal else S_k                 # :219
        for s0 in range(0, loop_end, BLOCK_N):
            s1 = min(s0 + BLOCK_N, S_k)
            cols = torch.arange(s0, s1)
            diag = not (q0 >= s0 + BLOCK_N - 1)                                 # :236-239
            P = _p_block(Qf[:, :, q0:q1], Kf[:, :, s0:s1], LSE[:, :, q0:q1], rows, cols,
                         scale, c, is_causal, diag)                             # :230-244
            dP = torch.matmul(dOf[:, :, q0:q1], Vf[:, :, s0:s1].transpose(-1, -2))  # :247
            dS = P * (dP - delta[:, :, q0:q1, None])                            # :250
            acc = acc + torch.matmul(_r16(dS, dt), Kf[:, :, s0:s1]) * scale     # :253
        # the reference stores dQ_acc.to(fp16) and the descriptor then casts to the run dtype (:256)
        dQ[:, :, q0:q1] = acc.to(dt)
    return dQ, delta


def dkv_blocked(Q, K, V, dO, LSE, delta, is_causal: bool, BLOCK_M: int = 64, BLOCK_N: int = 64,
                sm_scale: Optional[float] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """dK and dV, following code/_flash_attention_kernel_optimized.py:315-386."""
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    dt = Q.dtype
    scale = torch.tensor(1.0 / (D ** 0.5) if sm_scale is None else sm_scale, dtype=torch.float32)
    c = torch.tensor(LOG2_E, dtype=torch.float32)
    Qf, Kf, Vf, dOf = Q.float(), K.float(), V.float(), dO.float()
    dK = torch.zeros(B, H, S_k, D, dtype=dt)
    dV = torch.zeros(B, H, S_k, D, dtype=dt)
    for s0 in range(0, S_k, BLOCK_N):
        s1 = min(s0 + BLOCK_N, S_k)
        cols = torch.arange(s0, s1)
        dk = torch.zeros(B, H, s1 - s0, D)
        dv = torch.zeros(B, H, s1 - s0, D)
        loop_start = s0 if is_causal else 0                                     # :341
        for q0 in range(loop_start, S_q, BLOCK_M):
            q1 = min(q0 + BLOCK_M, S_q)
            rows = torch.arange(q0, q1)
            diag = not (q0 >= s0 + BLOCK_N - 1)                                 # :359-362
            P = _p_block(Qf[:, :, q0:q1], Kf[:, :, s0:s1], LSE[:, :, q0:q1], rows, cols,
                         scale, c, is_causal, diag)                             # :353-367
            dv = dv + torch.matmul(_r16(P.transpose(-1, -2), dt), dOf[:, :, q0:q1])   # :370
            dP = torch.matmul(dOf[:, :, q0:q1], Vf[:, :, s0:s1].transpose(-1, -2))    # :373
            dS = P * (dP - delta[:, :, q0:q1, None])                                  # :379
            dk = dk + torch.matmul(_r16(dS.transpose(-1, -2), dt), Qf[:, :, q0:q1]) * scale  # :382
        dK[:, :, s0:s1] = dk.to(dt)                                             # :385
        dV[:, :, s0:s1] = dv.to(dt)                                             # :386
    return dK, dV


def backward_blocked(Q, K, V, O, dO, LSE, is_causal, BLOCK_M=64, BLOCK_N=64, sm_scale=None):
    """dQ kernel then dKV kernel, as code/My_FlashAttention_optimized.py:62-128 launches them."""
    dQ, delta = dq_blocked(Q, K, V, O, dO, LSE, is_causal, BLOCK_M, BLOCK_N, sm_scale)
    dK, dV = dkv_blocked(Q, K, V, dO, LSE, delta, is_causal, BLOCK_M, BLOCK_N, sm_scale)
    return dQ, dK, dV, delta


def dropout_keep_mask(seed: int, B: int, H: int, S_q: int, S_k: int, dropout_p: float):
    """The keep mask and scale of the CUDA kernels' dropout (csrc/fa_ptx.cuh: mix32 / dropout_row_key / dropout_word /
    dropout_keep), restated with numpy uint32 arithmetic.  Returns (keep bool [B,H,S_q,S_k], scale)."""
    import numpy as np
    thresh = min(int(round(dropout_p * 256.0)), 255)
    M = np.uint32

    def mix32(x):
        x = x ^ (x >> M(16)); x = x * M(0x7FEB352D); x = x ^ (x >> M(15)); x = x * M(0x846CA68B); x = x ^ (x >> M(16))
        return x
    seed0, seed1 = M(seed & 0xFFFFFFFF), M((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        bh = np.arange(B * H, dtype=np.uint32)[:, None]
        q = np.arange(S_q, dtype=np.uint32)[None, :]
        row_key = mix32((bh * M(0x9E3779B1) + q) ^ seed0)                               # [BH, S_q]
        k = np.arange(S_k, dtype=np.uint32)
        word = mix32(row_key[:, :, None] ^ ((k >> M(2)) * M(0x85EBCA6B) + seed1)[None, None, :])   # [BH, S_q, S_k]
        byte = (word >> ((k & M(3)) * M(8))[None, None, :]) & M(0xFF)
    keep = torch.from_numpy((byte >= thresh).reshape(B, H, S_q, S_k))
    return keep, 256.0 / (256.0 - thresh)


def closed_form(Q, K, V, dO=None, is_causal: bool = False, sm_scale: Optional[float] = None,
                dtype: torch.dtype = torch.float64, q_offset: int = 0, k_offset: int = 0, row_ranges=None,
                keep_mask=None, keep_scale: float = 1.0):
    """Exact (materialised-scores) attention in ``dtype`` — no tiling, no 16-bit rounding.

    Forward: Phase_3.md:699-708 (LSE = logsumexp of masked, scaled scores).
    Backward: Phase_4.md:1251-1271 — dV = PᵀdO, dP = dO Vᵀ, delta = rowsum(dO∘O),
    dS = P∘(dP−delta), dQ = dS K·scale, dK = dSᵀ Q·scale.
    ``q_offset``/``k_offset`` shift the global row/col indices used by the causal mask
    (ring hops).  Fully-masked rows give O = 0, LSE = −inf.
    ``row_ranges`` = (lo, hi), int tensors [B, S_q]: query row i additionally sees only keys lo[b,i] <= j < hi[b,i]
    (packed variable-length sequences as in Phase_6.md:160-174, key padding, sliding windows).
    ``keep_mask`` [B,H,S_q,S_k] / ``keep_scale``: dropout on the attention probabilities (Phase_6.md:54-114):
    O = (P o keep * scale) V; the same mask scales dP in the backward; LSE and delta are unaffected.
    """
    D = Q.shape[-1]
    scale = 1.0 / math.sqrt(D) if sm_scale is None else sm_scale
    q, k, v = Q.to(dtype), K.to(dtype), V.to(dtype)
    S = torch.matmul(q, k.transpose(-1, -2)) * scale
    if is_causal:
        rows = torch.arange(Q.shape[2]) + q_offset
        cols = torch.arange(K.shape[2]) + k_offset
        S = S.masked_fill(~(rows[:, None] >= cols[None, :]), float("-inf"))
    if row_ranges is not None:
        lo, hi = (t.to(torch.int64).cpu() for t in row_ranges)
        cols = torch.arange(K.shape[2])
        keep = (cols[None, None, :] >= lo[:, :, None]) & (cols[None, None, :] < hi[:, :, None])      # [B, S_q, S_k]
        S = S.masked_fill(~keep[:, None], float("-inf"))
    LSE = torch.logsumexp(S, dim=-1)
    P = torch.exp(S - LSE[..., None])
    P = torch.nan_to_num(P, nan=0.0)           # fully masked rows: exp(-inf - -inf)
    Pd = P if keep_mask is None else P * keep_mask.to(dtype) * keep_scale
    O = torch.matmul(Pd, v)
    if dO is None:
        return O, LSE
    do = dO.to(dtype)
    dV = torch.matmul(Pd.transpose(-1, -2), do)
    dP = torch.matmul(do, v.transpose(-1, -2))
    if keep_mask is not None:
        dP = dP * keep_mask.to(dtype) * keep_scale
    delta = (do * O).sum(dim=-1)
    dS = P * (dP - delta[..., None])
    dQ = torch.matmul(dS, k) * scale
    dK = torch.matmul(dS.transpose(-1, -2), q) * scale
    return O, LSE, dQ, dK, dV


def sdpa_fp32(Q, K, V, dO=None, is_causal: bool = False, sm_scale: Optional[float] = None):
    """The reference's ground truth: fp32-upcast SDPA, MATH backend (Phase_3.md:244-252).

    Returns (O, dQ, dK, dV) in fp32 (grads only when ``dO`` is given).  LSE is not produced
    by SDPA; use ``closed_form`` or ``lse_bench`` for it.
    """
    import torch.nn.functional as F
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q = Q.detach().float().requires_grad_(dO is not None)
    k = K.detach().float().requires_grad_(dO is not None)
    v = V.detach().float().requires_grad_(dO is not None)
    with sdpa_kernel(SDPBackend.MATH):
        O = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0,
                                           is_causal=is_causal, scale=sm_scale)
    if dO is None:
        return O.detach()
    O.backward(dO.float())
    return O.detach(), q.grad, k.grad, v.grad


def sdpa_cpu_flash(Q, K, V, dO=None, is_causal: bool = False):
    """CPU flash backend of SDPA (no S materialisation) — the large-shape CPU baseline
    (BASELINE.md §4).  Returns (O, LSE) or (O, LSE, dQ, dK, dV); dtype follows the inputs."""
    q = Q.detach().requires_grad_(dO is not None)
    k = K.detach().requires_grad_(dO is not None)
    v = V.detach().requires_grad_(dO is not None)
    O, LSE = torch.ops.aten._scaled_dot_product_flash_attention_for_cpu(q, k, v, 0.0, is_causal)
    if dO is None:
        return O.detach(), LSE.detach()
    O.backward(dO.to(O.dtype))
    return O.detach(), LSE.detach(), q.grad, k.grad, v.grad


def lse_bench(Q, K, is_causal: bool, sm_scale: Optional[float] = None):
    """LSE oracle, Phase_3.md:699-708: logsumexp of masked QKᵀ·scale in fp32."""
    D = Q.shape[-1]
    scale = 1.0 / math.sqrt(D) if sm_scale is None else sm_scale
    S = torch.matmul(Q.float(), K.float().transpose(-1, -2)) * scale
    if is_causal:
        rows = torch.arange(Q.shape[2]); cols = torch.arange(K.shape[2])
        S = S.masked_fill(~(rows[:, None] >= cols[None, :]), float("-inf"))
    return torch.logsumexp(S, dim=-1)


def merge_partials(O_a, LSE_a, O_b, LSE_b):
    """(O, LSE) merge of two partial attentions over disjoint key sets:
    LSE = logaddexp(a, b); O = O_a·e^{LSE_a−LSE} + O_b·e^{LSE_b−LSE}.  Same algebra as the
    in-kernel correction (code/_flash_attention_kernel_optimized.py:106-117, :126).
    ``LSE = −inf`` partials are the identity."""
    LSE = torch.logaddexp(LSE_a, LSE_b)
    wa = torch.nan_to_num(torch.exp(LSE_a - LSE), nan=0.0)
    wb = torch.nan_to_num(torch.exp(LSE_b - LSE), nan=0.0)
    O = O_a.float() * wa[..., None] + O_b.float() * wb[..., None]
    return O, LSE


def make_inputs(B, H, S_q, S_k, D, dtype=torch.bfloat16, seed=0, with_dO=True):
    """Seeded synthetic inputs: N(0,1) drawn in fp32, then cast to the run dtype."""
    g = torch.Generator().manual_seed(seed)
    Q = torch.randn(B, H, S_q, D, generator=g).to(dtype)
    K = torch.randn(B, H, S_k, D, generator=g).to(dtype)
    V = torch.randn(B, H, S_k, D, generator=g).to(dtype)
    if not with_dO:
        return Q, K, V
    dO = torch.randn(B, H, S_q, D, generator=g).to(dtype)
    return Q, K, V, dO
