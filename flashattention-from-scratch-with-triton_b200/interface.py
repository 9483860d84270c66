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

import os

import torch

from . import _cabi

_DT = {torch.float16: _cabi.FA_DTYPE_FP16, torch.bfloat16: _cabi.FA_DTYPE_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def tma_compatible(t: torch.Tensor) -> bool:
    """True if a [B,H,S,D] tensor can be addressed by the kernels' 4-D tensor maps as it is: D contiguous,
    every other stride a multiple of 8 elements (16 bytes) for dims of extent > 1, 16-byte aligned base."""
    if t.ndim != 4 or t.stride(3) != 1 or t.data_ptr() % 16:
        return False
    return all(t.shape[i] == 1 or (t.stride(i) > 0 and t.stride(i) % 8 == 0) for i in range(3)) and \
        (t.shape[2] == 1 or t.stride(2) >= t.shape[3])


def as_kernel_layout(t: torch.Tensor) -> torch.Tensor:
    """The reference always calls .contiguous() (code/My_FlashAttention_optimized.py:138-140, :156); here a
    strided view (e.g. a [B,H,S,D] transpose of a [B,S,H,D] projection output) is used in place, and only
    layouts the TMA cannot express are copied."""
    return t if tma_compatible(t) else t.contiguous()


def _strides(*tensors):
    import ctypes
    arr = (ctypes.c_longlong * (3 * len(tensors)))()
    for i, t in enumerate(tensors):
        arr[3 * i], arr[3 * i + 1], arr[3 * i + 2] = t.stride(0), t.stride(1), t.stride(2)
    return arr


def flash_attention_forward(Q, K, V, is_causal, sm_scale=None):
    """Allocate O / LSE and launch the forward kernel (reference :14-60).

    Q: [B,H,S_q,D], K,V: [B,H,S_k,D] CUDA fp16/bf16, contiguous or TMA-compatible strided views.  Returns
    (O [B,H,S_q,D] in the input dtype and in Q's memory layout, LSE [B,H,S_q] fp32 = max + ln(sum exp) of the
    scaled scores)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    _, _, S_k, _ = K.shape
    O = torch.empty_like(Q)                               # contiguous Q -> contiguous O (reference :23); else Q's layout
    if not tma_compatible(O):
        O = torch.empty((B, H, S_q, D), dtype=Q.dtype, device=Q.device)
    LSE = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    st = _strides(Q, K, V, O)
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_fwd_strided(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), LSE.data_ptr(),
                                      B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                      float(sm_scale) if sm_scale is not None else 0.0, st, _stream(Q))
    _cabi.check("fa_sm100_fwd_strided", rc)
    return O, LSE


BWD_DELTA, BWD_DQ, BWD_DKV = 1, 2, 4


def flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, parts, sm_scale=None):
    """Launch a subset of the backward kernels into caller-provided outputs (per-kernel timing, ring hops)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    st = _strides(Q, K, V, O, dO, dQ, dK, dV)
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_bwd_strided(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                      LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                                      B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                      float(sm_scale) if sm_scale is not None else 0.0, st, _stream(Q), int(parts))
    _cabi.check("fa_sm100_bwd_strided", rc)


# Backward algorithm.  Head dim 64 defaults to the fused single-pass kernel (5 GEMMs and one exponential per score
# element; dQ summed over kv tiles by TMA reduce-add in fp32, so its low-order bits depend on scheduling — like
# PyTorch's own flash backward).  set_deterministic(True) or FA_SM100_DETERMINISTIC=1 selects the reference's
# atomic-free two-kernel structure (code/My_FlashAttention_optimized.py:111-126) everywhere; head dim 128 always uses it.
_deterministic = os.environ.get("FA_SM100_DETERMINISTIC", "0") not in ("", "0")


def set_deterministic(flag: bool) -> bool:
    """Select the bitwise-reproducible two-kernel backward for every head dim; returns the previous setting."""
    global _deterministic
    prev, _deterministic = _deterministic, bool(flag)
    return prev


def is_deterministic() -> bool:
    return _deterministic


def fused_backward_supported(Q) -> bool:
    return Q.shape[-1] == 64


BWD_FUSED, BWD_CONVERT = 8, 16


def flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, sm_scale=None, dq_acc=None, parts=0):
    """delta -> fused dK/dV/dQ kernel -> dQ conversion (head dim 64).  dq_acc: optional fp32 [B,H,S_q,D] workspace."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    if dq_acc is None:
        dq_acc = torch.empty((B, H, S_q, D), dtype=torch.float32, device=Q.device)
    assert dq_acc.dtype == torch.float32 and dq_acc.is_contiguous() and dq_acc.numel() == B * H * S_q * D
    st = _strides(Q, K, V, O, dO, dQ, dK, dV)
    with torch.cuda.device(Q.device):
        rc = lib.fa_sm100_bwd_fused(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                    LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                                    dq_acc.data_ptr(), B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                    float(sm_scale) if sm_scale is not None else 0.0, st, _stream(Q), int(parts))
    _cabi.check("fa_sm100_bwd_fused", rc)


def _empty_like_kernel(t):
    e = torch.empty_like(t)
    return e if tma_compatible(e) else torch.empty(t.shape, dtype=t.dtype, device=t.device)


def flash_attention_backward(Q, K, V, O, dO, LSE, is_causal, sm_scale=None):
    """Allocate dQ / dK / dV (+ fp32 delta) and launch the backward kernels (reference :62-128)."""
    B, H, S_q, D = Q.shape
    dQ, dK, dV = _empty_like_kernel(Q), _empty_like_kernel(K), _empty_like_kernel(V)      # reference :71-73
    delta = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    if fused_backward_supported(Q) and not _deterministic:
        flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, sm_scale)
    else:
        flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, BWD_DELTA | BWD_DQ | BWD_DKV, sm_scale)
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
        assert K.shape[:3] == V.shape[:3] and K.shape[0] == Q.shape[0] and Q.shape[1] % K.shape[1] == 0   # GQA/MQA: Hk | H
        Q_ = as_kernel_layout(Q); K_ = as_kernel_layout(K); V_ = as_kernel_layout(V)   # :138-140, without needless copies
        O, LSE = flash_attention_forward(Q_, K_, V_, is_causal, sm_scale)
        ctx.save_for_backward(Q_, K_, V_, O, LSE)                          # :145 (same set, same order)
        ctx.is_causal = is_causal                                          # :147
        ctx.sm_scale = sm_scale
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, LSE = ctx.saved_tensors                                # :154
        dO_ = as_kernel_layout(dO)                                         # :156
        dQ, dK, dV = flash_attention_backward(Q, K, V, O, dO_, LSE, ctx.is_causal, ctx.sm_scale)
        return dQ, dK, dV, None, None                                      # :166 (+None for sm_scale)


def flash_attention(Q, K, V, is_causal=False, *, sm_scale=None):
    """O = softmax(Q K^T * scale [+ causal mask]) V, differentiable w.r.t. Q, K, V (reference :169-170)."""
    return FlashAttentionFunction.apply(Q, K, V, is_causal, sm_scale)


attention = flash_attention


def flash_attention_bshd(q, k, v, is_causal=False, *, sm_scale=None):
    """Same operator for the [B, S, H, D] layout a fused QKV projection produces: zero-copy in and out
    (the kernels address the buffers through strided tensor maps).  Returns O as [B, S_q, H, D]."""
    O = flash_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), is_causal, sm_scale=sm_scale)
    return O.transpose(1, 2)


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
