"""The reference's PyTorch operator, re-hosted on hand-written sm_100a CUDA.

Mirror of /root/reference/code/My_FlashAttention_optimized.py:14-170 — same names, argument
order, dtypes, layouts, saved tensors and assertion behaviour — with the Triton launches
replaced by calls into libfa_sm100.so through the C ABI (include/fa_sm100.h):

    flash_attention(Q, K, V, is_causal=False)              reference :169-170
    FlashAttentionFunction.forward / .backward             reference :130-166
    flash_attention_forward(Q, K, V, is_causal)            reference :14-60
    flash_attention_backward(Q, K, V, O, dO, LSE, causal)  reference :62-128

Superset: keyword-only ``sm_scale`` (None -> the reference's hard-wired 1/sqrt(D), :56) and the
alias ``attention`` (BASELINE.json's spelling).
"""
from __future__ import annotations

import torch

from . import _cabi

_DT = {torch.float16: _cabi.FA_DTYPE_FP16, torch.bfloat16: _cabi.FA_DTYPE_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def flash_attention_forward(Q, K, V, is_causal, sm_scale=None):
    """Allocate O / LSE and launch the forward kernel (reference :14-60).

    Q: [B,H,S_q,D], K,V: [B,H,S_k,D] contiguous CUDA fp16/bf16.  Returns (O [B,H,S_q,D] in the
    input dtype, LSE [B,H,S_q] fp32 = max + ln(sum exp) of the scaled scores)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    _, _, S_k, _ = K.shape
    O = torch.empty((B, H, S_q, D), dtype=Q.dtype, device=Q.device)
    LSE = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_fwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), LSE.data_ptr(),
                              B, H, S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                              float(sm_scale) if sm_scale is not None else 0.0, _stream(Q))
    _cabi.check("fa_sm100_fwd", rc)
    return O, LSE


BWD_DELTA, BWD_DQ, BWD_DKV = 1, 2, 4


def flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, parts, sm_scale=None):
    """Launch a subset of the backward kernels into caller-provided outputs (per-kernel timing)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_bwd_parts(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                    LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                                    B, H, S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                    float(sm_scale) if sm_scale is not None else 0.0, _stream(Q), int(parts))
    _cabi.check("fa_sm100_bwd_parts", rc)


def flash_attention_backward(Q, K, V, O, dO, LSE, is_causal, sm_scale=None):
    """Allocate dQ / dK / dV (+ fp32 delta) and launch the backward kernels (reference :62-128)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    _, _, S_k, _ = K.shape
    dQ = torch.empty((B, H, S_q, D), dtype=Q.dtype, device=Q.device)
    dK = torch.empty((B, H, S_k, D), dtype=Q.dtype, device=Q.device)
    dV = torch.empty((B, H, S_k, D), dtype=Q.dtype, device=Q.device)
    delta = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_bwd(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                              LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                              B, H, S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                              float(sm_scale) if sm_scale is not None else 0.0, _stream(Q))
    _cabi.check("fa_sm100_bwd", rc)
    return dQ, dK, dV


class FlashAttentionFunction(torch.autograd.Function):
    """Same contract as the reference class (reference :130-166)."""

    @staticmethod
    def forward(ctx, Q, K, V, is_causal: bool, sm_scale=None):
        assert Q.is_cuda and K.is_cuda and V.is_cuda                       # :133
        assert Q.dtype in (torch.float16, torch.bfloat16)                  # :134
        assert Q.shape[-1] == K.shape[-1] == V.shape[-1]                   # :135
        assert Q.ndim == 4 and K.ndim == 4 and V.ndim == 4                 # :136
        assert K.dtype == Q.dtype and V.dtype == Q.dtype
        Q_ = Q.contiguous(); K_ = K.contiguous(); V_ = V.contiguous()      # :138-140
        O, LSE = flash_attention_forward(Q_, K_, V_, is_causal, sm_scale)
        ctx.save_for_backward(Q_, K_, V_, O, LSE)                          # :145 (same set, same order)
        ctx.is_causal = is_causal                                          # :147
        ctx.sm_scale = sm_scale
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, LSE = ctx.saved_tensors                                # :154
        dO_ = dO.contiguous()                                              # :156
        dQ, dK, dV = flash_attention_backward(Q, K, V, O, dO_, LSE, ctx.is_causal, ctx.sm_scale)
        return dQ, dK, dV, None, None                                      # :166 (+None for sm_scale)


def flash_attention(Q, K, V, is_causal=False, *, sm_scale=None):
    """O = softmax(Q K^T * scale [+ causal mask]) V, differentiable w.r.t. Q, K, V (reference :169-170)."""
    return FlashAttentionFunction.apply(Q, K, V, is_causal, sm_scale)


attention = flash_attention


def flash_attention_delta(O, dO):
    """delta = rowsum(dO * O) in fp32 (kernel :210-211) — the backward's preprocess, exposed for tests."""
    lib = _cabi.load()
    B, H, S_q, D = O.shape
    delta = torch.empty((B, H, S_q), dtype=torch.float32, device=O.device)
    with torch.cuda.device(O.device):
        rc = lib.fa_sm100_delta(O.data_ptr(), dO.data_ptr(), delta.data_ptr(), B, H, S_q, D, _DT[O.dtype], _stream(O))
    _cabi.check("fa_sm100_delta", rc)
    return delta


def merge_partial_(O_acc, LSE_acc, O_part, LSE_part, q_off=0):
    """In-place (O, LSE) merge of a partial attention over a disjoint key set (ring hops) into rows
    [q_off, q_off + S_part) of the fp32 accumulators O_acc [B,H,S_acc,D], LSE_acc [B,H,S_acc]."""
    lib = _cabi.load()
    B, H, S_q, D = O_part.shape
    S_acc = O_acc.shape[2]
    assert O_acc.dtype == torch.float32 and LSE_acc.dtype == torch.float32 and LSE_part.dtype == torch.float32
    assert O_acc.is_contiguous() and O_part.is_contiguous() and LSE_acc.is_contiguous() and LSE_part.is_contiguous()
    assert O_acc.shape[:2] == O_part.shape[:2] and O_acc.shape[3] == D and LSE_acc.shape == O_acc.shape[:3]
    with torch.cuda.device(O_part.device):
        rc = lib.fa_sm100_merge(O_acc.data_ptr(), LSE_acc.data_ptr(), O_part.data_ptr(), LSE_part.data_ptr(),
                                B, H, S_q, D, _DT[O_part.dtype], S_acc, int(q_off), _stream(O_part))
    _cabi.check("fa_sm100_merge", rc)
    return O_acc, LSE_acc
