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

import ctypes
import os

import torch

from . import _cabi

_DT = {torch.float16: _cabi.FA_DTYPE_FP16, torch.bfloat16: _cabi.FA_DTYPE_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


# Native host path (csrc/fa_torch.cpp -> _fa_torch.so): the same operator as FlashAttentionFunction below, as a C++ autograd node
# over the same C ABI.  flash_attention() uses it for the plain operator (contiguous inputs, no range mask, no dropout): a short
# step (C2: 0.2 ms of GPU work) otherwise spends a comparable time in Python autograd + ctypes marshalling, which is what limited
# the 8-rank run.  FA_SM100_HOST=python selects the Python class everywhere (A/B, debugging).  Neither is a
# fallback for the other's kernels: both launch libfa_sm100.so, and a missing _fa_torch.so is an ImportError.
_HOST_NATIVE = os.environ.get("FA_SM100_HOST", "native") != "python"
_native = None


def _load_native():
    global _native
    if _native is None:
        import importlib.util
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_fa_torch.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} not found: build it with `python flashattention-from-scratch-with-triton_b200/build.py` "
                              "(g++ against the torch headers), or set FA_SM100_HOST=python for the Python host path")
        _cabi.load()                                       # the library the extension links to (same file, loaded once)
        spec = importlib.util.spec_from_file_location("_fa_torch", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.set_deterministic(_deterministic)
        mod.set_fused128(globals().get("_fused128", False))
        _native = mod
    return _native


class _on_device:
    """torch.cuda.device(d) as a context only when d is not already current: the context manager costs ~10 us per call,
    comparable to the whole launch path, and a short step (C2: 0.2 ms on the GPU) must not become host-bound."""
    __slots__ = ("ctx",)

    def __init__(self, t: torch.Tensor):
        self.ctx = None if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def tma_compatible(t: torch.Tensor) -> bool:
    """True if a [B,H,S,D] tensor can be addressed by the kernels' 4-D tensor maps as it is: D contiguous,
    every other stride a multiple of 8 elements (16 bytes) for dims of extent > 1, 16-byte aligned base."""
    if t.ndim != 4 or t.data_ptr() % 16:
        return False
    if t.is_contiguous():                                  # fast path: D is a multiple of 8 for every supported head dim
        return t.shape[3] % 8 == 0
    if t.stride(3) != 1:
        return False
    return all(t.shape[i] == 1 or (t.stride(i) > 0 and t.stride(i) % 8 == 0) for i in range(3)) and \
        (t.shape[2] == 1 or t.stride(2) >= t.shape[3])


def as_kernel_layout(t: torch.Tensor) -> torch.Tensor:
    """The reference always calls .contiguous() (code/My_FlashAttention_optimized.py:138-140, :156); here a
    strided view (e.g. a [B,H,S,D] transpose of a [B,S,H,D] projection output) is used in place, and only
    layouts the TMA cannot express are copied."""
    if t.is_contiguous() and t.data_ptr() % 16 == 0 and t.ndim == 4:      # the common case, checked cheaply (host time per
        return t                                                          # step is comparable to a short step's GPU time)
    return t if tma_compatible(t) else t.contiguous()


def _strides(*tensors):
    if all(t.is_contiguous() for t in tensors):
        return None                                        # NULL = all contiguous [B,H,S,D] (include/fa_sm100.h)
    arr = (ctypes.c_longlong * (3 * len(tensors)))()
    for i, t in enumerate(tensors):
        arr[3 * i], arr[3 * i + 1], arr[3 * i + 2] = t.stride(0), t.stride(1), t.stride(2)
    return arr


class Ranges:
    """Range mask on top of (or instead of) causal: query row i of batch b sees keys [row_lo[b,i], row_hi[b,i]); the same mask seen
    from the key side is col_lo / col_hi [B, S_k].  int32 CUDA tensors, non-decreasing along the sequence.  Tiles outside the
    ranges are skipped by the kernels (include/fa_sm100.h, fa_sm100_*_ranges)."""

    def __init__(self, row_lo, row_hi, col_lo, col_hi):
        self.row_lo, self.row_hi, self.col_lo, self.col_hi = (t.to(torch.int32).contiguous() for t in (row_lo, row_hi, col_lo, col_hi))

    @staticmethod
    def from_cu_seqlens(cu_seqlens, total=None, device=None):
        """Packed self-attention (Phase_6.md:160-174): tokens [cu[s], cu[s+1]) form sequence s and attend only inside it.
        The ranges cover the cu[-1] packed tokens.  A packed BUFFER longer than that (total > cu[-1], a padded tail) is recorded
        in ``n_tokens`` / ``n_buffer``: flash_attention_varlen then hands the kernels views of the first cu[-1] tokens only, so the
        tail is never read (tensor-map out-of-bounds rows read as zero — it may hold anything, NaN included) and its outputs and
        gradients are zero.  Clamping tail tokens into the last sequence instead would make the row side and the key side of the
        mask disagree (the key side is what the dK/dV and fused kernels use)."""
        cu = torch.as_tensor(cu_seqlens, dtype=torch.int64, device=device)
        last = int(cu[-1])
        total = last if total is None else int(total)
        assert total >= last, f"packed buffer of {total} tokens is shorter than cu_seqlens[-1] = {last}"
        pos = torch.arange(last, device=cu.device)
        sid = torch.searchsorted(cu[1:].contiguous(), pos, right=True).clamp_(max=cu.numel() - 2)
        lo, hi = cu[sid][None], cu[sid + 1][None]
        r = Ranges(lo, hi, lo, hi)
        r.n_tokens, r.n_buffer = last, total
        return r

    def validate(self, is_causal=False):
        """Check the documented preconditions (include/fa_sm100.h): all four arrays non-decreasing along the sequence,
        lo <= hi, and the row side and the key side describing the same mask (query i sees key j  <=>  key j is seen by
        query i).  O(B*Sq*Sk) memory: a debugging aid for hand-built masks, not part of the hot path.  Rows with an empty
        range are legal (O = 0, LSE = -inf, zero gradients)."""
        rl, rh, cl, ch = (t.long() for t in (self.row_lo, self.row_hi, self.col_lo, self.col_hi))
        for name, t in (("row_lo", rl), ("row_hi", rh), ("col_lo", cl), ("col_hi", ch)):
            if t.shape[1] > 1 and bool((t[:, 1:] < t[:, :-1]).any()):
                raise ValueError(f"Ranges.{name} must be non-decreasing along the sequence")
        if bool((rl > rh).any()) or bool((cl > ch).any()):
            raise ValueError("Ranges: lo must not exceed hi")
        Sq, Sk = rl.shape[1], cl.shape[1]
        i = torch.arange(Sq, device=rl.device)[None, :, None]; j = torch.arange(Sk, device=rl.device)[None, None, :]
        row_view = (j >= rl[:, :, None]) & (j < rh[:, :, None].clamp(max=Sk))
        col_view = (i >= cl[:, None, :]) & (i < ch[:, None, :].clamp(max=Sq))
        if is_causal:
            row_view &= (j <= i); col_view &= (j <= i)
        if not torch.equal(row_view, col_view):
            raise ValueError("Ranges: row_lo/row_hi and col_lo/col_hi describe different masks")
        return self

    @staticmethod
    def from_key_padding(seqlens_k, S_q, S_k, device=None):
        """Padded batch: batch b has seqlens_k[b] valid keys (>= 1); padded keys receive zero gradient."""
        n = torch.as_tensor(seqlens_k, dtype=torch.int64, device=device)
        B = n.numel()
        row_lo = torch.zeros(B, S_q, dtype=torch.int64, device=n.device)
        row_hi = n[:, None].expand(B, S_q)
        valid = torch.arange(S_k, device=n.device)[None, :] < n[:, None]
        col_lo = torch.where(valid, 0, S_q)
        col_hi = torch.full((B, S_k), S_q, dtype=torch.int64, device=n.device)
        return Ranges(row_lo, row_hi, col_lo, col_hi)

    @staticmethod
    def sliding_window(B, S, window, device=None):
        """Causal sliding window: query i sees keys (i - window, i].  Use with is_causal=True."""
        i = torch.arange(S, device=device)
        row_lo = (i - window + 1).clamp_(min=0)[None].expand(B, S)
        row_hi = torch.full((B, S), S, dtype=torch.int64, device=device)
        col_lo = i[None].expand(B, S)
        col_hi = (i + window).clamp_(max=S)[None].expand(B, S)
        return Ranges(row_lo, row_hi, col_lo, col_hi)

    def row_ranges(self):
        return self.row_lo, self.row_hi


def _options(ranges, dropout_p, dropout_seed, backward=False):
    """ctypes fa_sm100_options (or None for the plain operator)."""
    if ranges is None and not dropout_p:
        return None
    o = _cabi.Options()
    if ranges is not None:
        o.row_lo, o.row_hi = ranges.row_lo.data_ptr(), ranges.row_hi.data_ptr()
        if backward:
            o.col_lo, o.col_hi = ranges.col_lo.data_ptr(), ranges.col_hi.data_ptr()
    o.dropout_p = float(dropout_p or 0.0)
    o.dropout_seed = int(dropout_seed or 0) & 0xFFFFFFFFFFFFFFFF
    return ctypes.byref(o)


def flash_attention_forward(Q, K, V, is_causal, sm_scale=None, ranges=None, dropout_p=0.0, dropout_seed=0):
    """Allocate O / LSE and launch the forward kernel (reference :14-60).

    Q: [B,H,S_q,D], K,V: [B,H,S_k,D] CUDA fp16/bf16, contiguous or TMA-compatible strided views.  Returns
    (O [B,H,S_q,D] in the input dtype and in Q's memory layout, LSE [B,H,S_q] fp32 = max + ln(sum exp) of the
    scaled scores)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    _, _, S_k, _ = K.shape
    O = torch.empty_like(Q)                               # contiguous Q -> contiguous O (reference :23); else Q's layout
    if not O.is_contiguous() and not tma_compatible(O):
        O = torch.empty((B, H, S_q, D), dtype=Q.dtype, device=Q.device)
    LSE = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    st = _strides(Q, K, V, O)
    with _on_device(Q):
        if ranges is not None:
            assert ranges.row_lo.shape == (B, S_q) and ranges.row_hi.shape == (B, S_q) and ranges.row_lo.device == Q.device
        rc = lib.fa_sm100_fwd_opt(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), LSE.data_ptr(),
                                  B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                  float(sm_scale) if sm_scale is not None else 0.0, st,
                                  _options(ranges, dropout_p, dropout_seed), _stream(Q))
    _cabi.check("fa_sm100_fwd_opt", rc)
    return O, LSE


BWD_DELTA, BWD_DQ, BWD_DKV = 1, 2, 4


def flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, parts, sm_scale=None, ranges=None,
                                   dropout_p=0.0, dropout_seed=0):
    """Launch a subset of the backward kernels into caller-provided outputs (per-kernel timing, ring hops)."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    st = _strides(Q, K, V, O, dO, dQ, dK, dV)
    with _on_device(Q):
        if ranges is not None:
            assert ranges.col_lo.shape == (B, S_k) and ranges.col_hi.shape == (B, S_k) and ranges.row_lo.shape == (B, S_q)
        rc = lib.fa_sm100_bwd_opt(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                  LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                                  B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                  float(sm_scale) if sm_scale is not None else 0.0, st,
                                  _options(ranges, dropout_p, dropout_seed, backward=True), _stream(Q), int(parts))
    _cabi.check("fa_sm100_bwd_opt", rc)


# Backward algorithm.  Head dim 64 defaults to the fused single-pass kernel (5 GEMMs and one exponential per score
# element; dQ summed over kv tiles by TMA reduce-add in fp32, so its low-order bits depend on scheduling — like
# PyTorch's own flash backward).  set_deterministic(True) or FA_SM100_DETERMINISTIC=1 selects the reference's
# atomic-free two-kernel structure (code/My_FlashAttention_optimized.py:111-126) everywhere; head dim 128 uses it unless
# set_fused128(True) / FA_SM100_FUSED128=1 (measured at parity there).
_deterministic = os.environ.get("FA_SM100_DETERMINISTIC", "0") not in ("", "0")


def set_deterministic(flag: bool) -> bool:
    """Select the bitwise-reproducible two-kernel backward for every head dim; returns the previous setting."""
    global _deterministic
    prev, _deterministic = _deterministic, bool(flag)
    if _native is not None:
        _native.set_deterministic(_deterministic)
    return prev


def set_shared_sms(flag: bool) -> bool:
    """Tell the kernels that their launches share the GPU with other kernels (NCCL transfers of the sequence-parallel paths):
    persistent CTAs then draw even their first work item from the counter (include/fa_sm100.h).  Returns the previous setting."""
    return bool(_cabi.load().fa_sm100_set_shared_sms(int(bool(flag))))


def is_deterministic() -> bool:
    return _deterministic


# Head dim 128: fused kernel (csrc/fa_bwd_fused128.cuh) for the plain operator (no range masks, no dropout).
_fused128 = os.environ.get("FA_SM100_FUSED128", "0") not in ("", "0")


def set_fused128(flag: bool) -> bool:
    """Use the fused single-pass backward at head dim 128 too (plain operator only); returns the previous setting."""
    global _fused128
    prev, _fused128 = _fused128, bool(flag)
    if _native is not None and hasattr(_native, "set_fused128"):
        _native.set_fused128(_fused128)
    return prev


def fused_backward_supported(Q, ranges=None, dropout_p=0.0) -> bool:
    D = Q.shape[-1]
    return D == 64 or (D == 128 and _fused128 and ranges is None and not dropout_p)


BWD_FUSED, BWD_CONVERT = 8, 16


def flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, sm_scale=None, dq_acc=None, parts=0,
                                   ranges=None, dropout_p=0.0, dropout_seed=0):
    """delta -> fused dK/dV/dQ kernel -> dQ conversion (head dim 64, or 128 for the plain operator: no ranges / dropout there).
    dq_acc: optional fp32 [B,H,S_q,D] workspace."""
    lib = _cabi.load()
    B, H, S_q, D = Q.shape
    S_k = K.shape[2]
    if dq_acc is None:
        dq_acc = torch.empty((B, H, S_q, D), dtype=torch.float32, device=Q.device)
    assert dq_acc.dtype == torch.float32 and dq_acc.is_contiguous() and dq_acc.numel() == B * H * S_q * D
    st = _strides(Q, K, V, O, dO, dQ, dK, dV)
    with _on_device(Q):
        rc = lib.fa_sm100_bwd_fused_opt(Q.data_ptr(), K.data_ptr(), V.data_ptr(), O.data_ptr(), dO.data_ptr(),
                                        LSE.data_ptr(), dQ.data_ptr(), dK.data_ptr(), dV.data_ptr(), delta.data_ptr(),
                                        dq_acc.data_ptr(), B, H, K.shape[1], S_q, S_k, D, _DT[Q.dtype], int(bool(is_causal)),
                                        float(sm_scale) if sm_scale is not None else 0.0, st,
                                        _options(ranges, dropout_p, dropout_seed, backward=True), _stream(Q), int(parts))
    _cabi.check("fa_sm100_bwd_fused_opt", rc)


def _empty_like_kernel(t):
    e = torch.empty_like(t)
    return e if e.is_contiguous() or tma_compatible(e) else torch.empty(t.shape, dtype=t.dtype, device=t.device)


def flash_attention_backward(Q, K, V, O, dO, LSE, is_causal, sm_scale=None, ranges=None, dropout_p=0.0, dropout_seed=0):
    """Allocate dQ / dK / dV (+ fp32 delta) and launch the backward kernels (reference :62-128)."""
    B, H, S_q, D = Q.shape
    dQ, dK, dV = _empty_like_kernel(Q), _empty_like_kernel(K), _empty_like_kernel(V)      # reference :71-73
    delta = torch.empty((B, H, S_q), dtype=torch.float32, device=Q.device)
    if fused_backward_supported(Q, ranges, dropout_p) and not _deterministic:
        flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, sm_scale, None, 0, ranges, dropout_p,
                                       dropout_seed)
    else:
        flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, is_causal, BWD_DELTA | BWD_DQ | BWD_DKV, sm_scale,
                                       ranges, dropout_p, dropout_seed)
    return dQ, dK, dV


class FlashAttentionFunction(torch.autograd.Function):
    """Same contract as the reference class (reference :130-166)."""

    @staticmethod
    def forward(ctx, Q, K, V, is_causal: bool, sm_scale=None, ranges=None, dropout_p=0.0, dropout_seed=0):
        assert Q.is_cuda and K.is_cuda and V.is_cuda                       # :133
        assert Q.dtype in (torch.float16, torch.bfloat16)                  # :134
        assert Q.shape[-1] == K.shape[-1] == V.shape[-1]                   # :135
        assert Q.ndim == 4 and K.ndim == 4 and V.ndim == 4                 # :136
        assert K.dtype == Q.dtype and V.dtype == Q.dtype
        assert K.shape[:3] == V.shape[:3] and K.shape[0] == Q.shape[0] and Q.shape[1] % K.shape[1] == 0   # GQA/MQA: Hk | H
        Q_ = as_kernel_layout(Q); K_ = as_kernel_layout(K); V_ = as_kernel_layout(V)   # :138-140, without needless copies
        O, LSE = flash_attention_forward(Q_, K_, V_, is_causal, sm_scale, ranges, dropout_p, dropout_seed)
        ctx.save_for_backward(Q_, K_, V_, O, LSE)                          # :145 (same set, same order)
        ctx.is_causal = is_causal                                          # :147
        ctx.sm_scale = sm_scale
        ctx.ranges = ranges
        ctx.dropout = (dropout_p, dropout_seed)           # the backward regenerates the keep mask from the same seed
        return O

    @staticmethod
    def backward(ctx, dO):
        Q, K, V, O, LSE = ctx.saved_tensors                                # :154
        dO_ = as_kernel_layout(dO)                                         # :156
        dQ, dK, dV = flash_attention_backward(Q, K, V, O, dO_, LSE, ctx.is_causal, ctx.sm_scale, ctx.ranges, *ctx.dropout)
        return dQ, dK, dV, None, None, None, None, None                    # :166 (+None for every added argument)


def flash_attention(Q, K, V, is_causal=False, *, sm_scale=None, ranges=None, dropout_p=0.0, dropout_seed=None):
    """O = softmax(Q K^T * scale [+ causal mask]) V, differentiable w.r.t. Q, K, V (reference :169-170).
    ``ranges`` (a Ranges object) adds a per-row key-range mask: packed sequences, key padding, sliding windows.
    ``dropout_p`` > 0 drops attention probabilities (quantised to 1/256, kept ones scaled by 1/(1-p)); the mask is a pure
    function of ``dropout_seed`` (drawn from torch's CPU generator when None) and the element's coordinates."""
    if _HOST_NATIVE and ranges is None and not dropout_p:
        # the reference's asserts (:133-136), then the C++ autograd node when the layout needs no tensor-map strides
        assert Q.is_cuda and K.is_cuda and V.is_cuda
        assert Q.dtype in (torch.float16, torch.bfloat16)
        assert Q.shape[-1] == K.shape[-1] == V.shape[-1]
        assert Q.ndim == 4 and K.ndim == 4 and V.ndim == 4
        assert K.dtype == Q.dtype and V.dtype == Q.dtype
        assert K.shape[:3] == V.shape[:3] and K.shape[0] == Q.shape[0] and Q.shape[1] % K.shape[1] == 0
        if Q.is_contiguous() and K.is_contiguous() and V.is_contiguous():
            return (_native or _load_native()).flash_attention(Q, K, V, bool(is_causal), float(sm_scale) if sm_scale is not None else 0.0)
    if dropout_p and dropout_seed is None:
        dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return FlashAttentionFunction.apply(Q, K, V, is_causal, sm_scale, ranges, float(dropout_p or 0.0), dropout_seed or 0)


def flash_attention_varlen(q, k, v, cu_seqlens, is_causal=False, *, sm_scale=None, ranges=None):
    """Variable-length self-attention over PACKED sequences (the tutorial's own next step, Phase_6.md:160-174):
    q [total, H, D], k / v [total, Hk, D]; tokens cu_seqlens[s] .. cu_seqlens[s+1]-1 form sequence s.  Zero-copy: the packed
    buffers are addressed as one batch of length `total` through strided tensor maps, the block-diagonal mask is a Ranges
    object and the kernels skip every tile that lies outside a sequence.  Returns O [total, H, D].
    Building the Ranges from cu_seqlens costs a handful of tiny launches: pass ``ranges=Ranges.from_cu_seqlens(...)`` to reuse
    one object across the layers of a model (cu_seqlens is then ignored)."""
    total = q.shape[0]
    if ranges is None:
        ranges = Ranges.from_cu_seqlens(cu_seqlens, total, device=q.device)
    n = getattr(ranges, "n_tokens", total)                 # tokens the ranges describe; a longer buffer has a padded tail
    assert ranges.row_lo.shape[1] == n <= total, "ranges do not match the packed buffer"
    O = flash_attention(q[:n].transpose(0, 1)[None], k[:n].transpose(0, 1)[None], v[:n].transpose(0, 1)[None], is_causal,
                        sm_scale=sm_scale, ranges=ranges)
    O = O[0].transpose(0, 1)
    if n < total:                                          # pad rows: zeros (and zero gradients, through the view's autograd)
        O = torch.cat([O, O.new_zeros((total - n,) + tuple(O.shape[1:]))])
    return O


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
    with _on_device(O):
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
    with _on_device(O_part):
        rc = lib.fa_sm100_merge(O_acc.data_ptr(), LSE_acc.data_ptr(), O_part.data_ptr(), LSE_part.data_ptr(),
                                B, H, S_q, D, _DT[O_part.dtype], S_acc, int(q_off), _stream(O_part))
    _cabi.check("fa_sm100_merge", rc)
    return O_acc, LSE_acc
