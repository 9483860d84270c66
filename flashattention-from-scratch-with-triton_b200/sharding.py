"""Multi-GPU execution of the operator (no counterpart in the reference, which is single-GPU;
BASELINE.json north_star (3)).  One process per GPU, ``torch.distributed`` for plumbing.

* batch x head sharding (config C4): every (b, h) pair is an independent attention problem (the reference's
  grid axis 1, code/My_FlashAttention_optimized.py:53), so ranks take disjoint slices and run the
  single-GPU operator.  NO collective on the data path.
* sequence-sharded ring (config C5, causal long context): the sequence is cut into 2P chunks and rank r owns
  chunks r and 2P-1-r ("zigzag"), which balances causal work.  K/V blocks travel around the ring with
  point-to-point send/recv (NCCL over NVLink on GPUs), double-buffered under compute.  Each hop is one of
      hop from self      : local causal attention over [chunk r, chunk 2P-1-r]
      hop from rank o < r : all local queries  x  first half of the visiting K/V (non-causal)
      hop from rank o > r : second half of the local queries  x  all of the visiting K/V (non-causal)
  so the single-GPU kernels need nothing but their causal / non-causal modes with S_q != S_k, and every hop
  costs the same.  Partials are merged with (O, LSE) log-sum-exp algebra (fa_sm100_merge).  The backward
  sends dK/dV accumulators around the ring with their K/V block; they arrive home after P hops.

The local math is injected through an ``ops`` object so that the schedule can be verified on CPU ranks
(gloo) in tests/ with the CPU oracle as the local kernel; the default ``CudaOps`` is the sm_100a library.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

# ------------------------------------------------------------------------------------------------
# batch x head sharding
# ------------------------------------------------------------------------------------------------


def partition_batch_heads(B: int, H: int, world: int) -> List[Tuple[int, int]]:
    """Balanced contiguous ranges [lo, hi) over the flattened (b*H + h) index, one per rank."""
    n = B * H
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi)); lo = hi
    return out


def local_shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's slice of a [B,H,S,D] tensor as [1, n_local, S, D] (a view when the tensor is
    contiguous: (b,h) pairs are contiguous blocks of S*D elements)."""
    B, H, S, D = t.shape
    lo, hi = partition_batch_heads(B, H, world)[rank]
    return t.reshape(1, B * H, S, D)[:, lo:hi]


def sharded_flash_attention(Q, K, V, is_causal=False, rank=0, world=1, attn=None):
    """Run the operator on this rank's (b,h) slice of globally-shaped inputs.  Returns the local output
    [1, n_local, S_q, D]; no communication happens."""
    if attn is None:
        from .interface import flash_attention as attn
    q, k, v = (local_shard(t, rank, world) for t in (Q, K, V))
    return attn(q, k, v, is_causal)


# ------------------------------------------------------------------------------------------------
# zigzag sequence partition
# ------------------------------------------------------------------------------------------------


def zigzag_chunks(rank: int, world: int) -> Tuple[int, int]:
    return rank, 2 * world - 1 - rank


def zigzag_split(t: torch.Tensor, rank: int, world: int, dim: int = 2) -> torch.Tensor:
    """Local part of a globally-ordered sequence tensor: chunks (rank, 2P-1-rank) concatenated."""
    S = t.shape[dim]
    assert S % (2 * world) == 0, "sequence length must divide into 2*world chunks"
    c = S // (2 * world)
    a, b = zigzag_chunks(rank, world)
    return torch.cat([t.narrow(dim, a * c, c), t.narrow(dim, b * c, c)], dim=dim).contiguous()


def zigzag_merge(parts: List[torch.Tensor], dim: int = 2) -> torch.Tensor:
    """Inverse of zigzag_split over all ranks' local tensors."""
    world = len(parts)
    c = parts[0].shape[dim] // 2
    chunks = [None] * (2 * world)
    for r, p in enumerate(parts):
        a, b = zigzag_chunks(r, world)
        chunks[a] = p.narrow(dim, 0, c); chunks[b] = p.narrow(dim, c, c)
    return torch.cat(chunks, dim=dim)


# ------------------------------------------------------------------------------------------------
# local ops
# ------------------------------------------------------------------------------------------------


class CudaOps:
    """Local kernels = the sm_100a library through the C ABI."""

    def fwd(self, q, k, v, causal):
        from .interface import flash_attention_forward
        return flash_attention_forward(q, k, v, causal)

    def merge_(self, O_acc, LSE_acc, O_part, LSE_part, q_off):
        from .interface import merge_partial_
        merge_partial_(O_acc, LSE_acc, O_part, LSE_part, q_off)

    def delta(self, O, dO):
        from .interface import flash_attention_delta
        return flash_attention_delta(O, dO)

    def bwd(self, q, k, v, o, do, lse, delta, causal):
        """dq, dk, dv of one hop given the GLOBAL lse / delta of the local query rows."""
        from .interface import BWD_DKV, BWD_DQ, flash_attention_backward_parts
        dq = torch.empty_like(q); dk = torch.empty_like(k); dv = torch.empty_like(v)
        flash_attention_backward_parts(q, k, v, o, do, lse, dq, dk, dv, delta, causal, BWD_DQ | BWD_DKV)
        return dq, dk, dv


def _exchange(send: List[torch.Tensor], recv: List[torch.Tensor], group, rank: int, world: int):
    """Post send-to-next / recv-from-prev for a list of tensors; returns the requests."""
    import torch.distributed as dist
    nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    ops = []
    for s, r in zip(send, recv):
        ops.append(dist.P2POp(dist.isend, s, nxt, group))
        ops.append(dist.P2POp(dist.irecv, r, prv, group))
    return dist.batch_isend_irecv(ops)


# ------------------------------------------------------------------------------------------------
# ring forward / backward
# ------------------------------------------------------------------------------------------------


def ring_attention_forward(q, k, v, group=None, ops=None):
    """Causal attention over the global sequence; q,k,v are this rank's zigzag-local [B,H,2c,D] tensors.
    Returns (O [B,H,2c,D] in q.dtype, LSE [B,H,2c] fp32) for the local rows."""
    import torch.distributed as dist
    ops = ops or CudaOps()
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    B, H, S2, D = q.shape
    c = S2 // 2
    O_acc = torch.zeros(B, H, S2, D, dtype=torch.float32, device=q.device)
    L_acc = torch.full((B, H, S2), float("-inf"), dtype=torch.float32, device=q.device)
    q_hi = q[:, :, c:].contiguous()
    kv = [k.contiguous(), v.contiguous()]
    for s in range(world):
        reqs, nxt = [], None
        if s + 1 < world:                                     # prefetch the next visiting block under this hop's compute
            nxt = [torch.empty_like(kv[0]), torch.empty_like(kv[1])]
            reqs = _exchange(kv, nxt, group, rank, world)
        o = (rank - s) % world                                # owner of the visiting K/V block
        if o == rank:
            Op, Lp = ops.fwd(q, kv[0], kv[1], True)
            ops.merge_(O_acc, L_acc, Op, Lp, 0)
        elif o < rank:
            Op, Lp = ops.fwd(q, kv[0][:, :, :c].contiguous(), kv[1][:, :, :c].contiguous(), False)
            ops.merge_(O_acc, L_acc, Op, Lp, 0)
        else:
            Op, Lp = ops.fwd(q_hi, kv[0], kv[1], False)
            ops.merge_(O_acc, L_acc, Op, Lp, c)
        for r in reqs:
            r.wait()
        if nxt is not None:
            kv = nxt
    return O_acc.to(q.dtype), L_acc


def ring_attention_backward(q, k, v, O, dO, LSE, group=None, ops=None):
    """Gradients for ring_attention_forward.  Returns (dq, dk, dv) for the local rows, in q.dtype.

    Two rings run under the compute: the K/V block of the next hop is prefetched while the current hop's
    kernels run, and the fp32 dK/dV accumulators of the block that just left are in flight to the next rank
    while this rank already computes its contribution to the following block; they are added on arrival."""
    import torch.distributed as dist
    ops = ops or CudaOps()
    world = dist.get_world_size(group); rank = dist.get_rank(group)
    B, H, S2, D = q.shape
    c = S2 // 2
    f32 = dict(dtype=torch.float32, device=q.device)
    delta = ops.delta(O.contiguous(), dO.contiguous())        # global: uses the final O of the local rows
    dq_acc = torch.zeros(B, H, S2, D, **f32)
    hi = lambda t: t[:, :, c:].contiguous()
    q_hi, O_hi, dO_hi, L_hi, d_hi = hi(q), hi(O), hi(dO), hi(LSE), hi(delta)
    kv = [k.contiguous(), v.contiguous()]
    dkv = [torch.zeros(B, H, S2, D, **f32), torch.zeros(B, H, S2, D, **f32)]   # accumulators of the visiting block
    dkv_reqs, dkv_in = [], None
    for s in range(world):
        kv_reqs, nxt = [], None
        if s + 1 < world:
            nxt = [torch.empty_like(kv[0]), torch.empty_like(kv[1])]
            kv_reqs = _exchange(kv, nxt, group, rank, world)
        o = (rank - s) % world
        if o == rank:
            dq, dk, dv = ops.bwd(q, kv[0], kv[1], O, dO, LSE, delta, True)
            part = (slice(None), slice(None), slice(0, S2))
            dq_acc.add_(dq)
        elif o < rank:
            dq, dk, dv = ops.bwd(q, kv[0][:, :, :c].contiguous(), kv[1][:, :, :c].contiguous(), O, dO, LSE, delta, False)
            part = (slice(None), slice(None), slice(0, c))
            dq_acc.add_(dq)
        else:
            dq, dk, dv = ops.bwd(q_hi, kv[0], kv[1], O_hi, dO_hi, L_hi, d_hi, False)
            part = (slice(None), slice(None), slice(0, S2))
            dq_acc[:, :, c:].add_(dq)
        if s > 0:                                             # accumulators of this block, sent by the previous rank
            for r in dkv_reqs:
                r.wait()
            dkv = dkv_in
        dkv[0][part].add_(dk); dkv[1][part].add_(dv)
        # pass them on with their block (after the last hop they arrive back at the block's owner)
        dkv_in = [torch.empty_like(dkv[0]), torch.empty_like(dkv[1])]
        dkv_reqs = _exchange(dkv, dkv_in, group, rank, world)
        for r in kv_reqs:
            r.wait()
        if nxt is not None:
            kv = nxt
    for r in dkv_reqs:
        r.wait()
    return dq_acc.to(q.dtype), dkv_in[0].to(q.dtype), dkv_in[1].to(q.dtype)


class RingFlashAttentionFunction(torch.autograd.Function):
    """Autograd wrapper with the same saved-tensor contract as FlashAttentionFunction (Q,K,V,O,LSE)."""

    @staticmethod
    def forward(ctx, q, k, v, group=None):
        assert q.dtype in (torch.float16, torch.bfloat16) and q.ndim == 4
        q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
        O, LSE = ring_attention_forward(q_, k_, v_, group)
        ctx.save_for_backward(q_, k_, v_, O, LSE)
        ctx.group = group
        return O

    @staticmethod
    def backward(ctx, dO):
        q, k, v, O, LSE = ctx.saved_tensors
        dq, dk, dv = ring_attention_backward(q, k, v, O, dO.contiguous(), LSE, ctx.group)
        return dq, dk, dv, None


def ring_flash_attention(q, k, v, group=None):
    """Causal attention over a sequence sharded zigzag-wise across the ranks of `group`."""
    return RingFlashAttentionFunction.apply(q, k, v, group)
