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


class DistComm:
    """Ring neighbours over ``torch.distributed`` point-to-point ops (NCCL over NVLink on GPUs, gloo on CPU ranks)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group); self.rank = dist.get_rank(group)
        g = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        self.nxt, self.prv = g((self.rank + 1) % self.world), g((self.rank - 1) % self.world)

    def exchange(self, send: List[torch.Tensor], recv: List[torch.Tensor]):
        """Post send-to-next / recv-from-prev for a list of tensors; returns an object with .wait()."""
        import torch.distributed as dist
        ops = []
        for s, r in zip(send, recv):
            ops.append(dist.P2POp(dist.isend, s, self.nxt, self.group))
            ops.append(dist.P2POp(dist.irecv, r, self.prv, self.group))
        return _Pending(dist.batch_isend_irecv(ops))


class _Pending:
    def __init__(self, reqs):
        self.reqs = reqs

    def wait(self):
        for r in self.reqs:
            r.wait()


def _parts(B: int, Hk: int, G: int, n: int):
    """Cut one hop's work into <= n independent sub-launches: along the batch if B > 1, else along the K/V heads (slices of a
    contiguous [B,H,S,D] tensor that are contiguous themselves).  Yields (q-side index, kv-side index) tuples."""
    if B > 1:
        n = max(1, min(n, B)); per = -(-B // n)
        return [((slice(b, min(b + per, B)),), (slice(b, min(b + per, B)),)) for b in range(0, B, per)]
    n = max(1, min(n, Hk)); per = -(-Hk // n)
    return [((slice(None), slice(h * G, min(h + per, Hk) * G)), (slice(None), slice(h, min(h + per, Hk)))) for h in range(0, Hk, per)]


def _mark(timeline, t, name):
    """Optional per-hop timeline: CUDA events on the current stream (GPU tensors only)."""
    if timeline is not None and t.is_cuda:
        ev = torch.cuda.Event(enable_timing=True); ev.record()
        timeline.append((name, ev))


# ------------------------------------------------------------------------------------------------
# ring forward / backward
# ------------------------------------------------------------------------------------------------
# Why a hop is cut into sub-launches (`splits`): the attention kernels are persistent, one CTA per SM using the SM's whole
# register file, so while one of them runs NOTHING else can be scheduled on the GPU — including NCCL's send/recv kernels.  A hop
# issued as one launch therefore serialises "transfer" and "compute" no matter which stream they are on.  Cut into a few
# launches (by K/V head or batch: independent problems), the SMs change hands at every boundary and the pending NCCL kernel gets
# its CTAs; the dynamic tile scheduler of the next launch simply runs on the SMs that are left.


def ring_attention_forward(q, k, v, group=None, ops=None, comm=None, splits=None, timeline=None):
    """Causal attention over the global sequence; q,k,v are this rank's zigzag-local [B,H,2c,D] tensors.
    Returns (O [B,H,2c,D] in q.dtype, LSE [B,H,2c] fp32) for the local rows."""
    ops = ops or CudaOps()
    comm = comm or DistComm(group)
    world, rank = comm.world, comm.rank
    B, H, S2, D = q.shape
    Hk = k.shape[1]; G = H // Hk
    c = S2 // 2
    parts = _parts(B, Hk, G, splits if splits is not None else (4 if world > 1 else 1))
    O_acc = torch.empty(B, H, S2, D, dtype=torch.float32, device=q.device)
    L_acc = torch.empty((B, H, S2), dtype=torch.float32, device=q.device)
    q = q.contiguous()
    cur = [k.contiguous(), v.contiguous()]
    # two receive buffers per tensor, allocated once: hop s computes on `cur` while the block of hop s+1 lands in bufs[s % 2]
    bufs = [[torch.empty_like(cur[0]), torch.empty_like(cur[1])] for _ in range(min(2, world - 1))]
    for s in range(world):
        _mark(timeline, q, f"fwd{s}:start")
        pending, nxt = None, None
        if s + 1 < world:                                     # prefetch the next visiting block under this hop's compute
            nxt = bufs[s % 2]
            pending = comm.exchange(cur, nxt)
        o = (rank - s) % world                                # owner of the visiting K/V block
        for qi, ki in parts:
            if o == rank:                                     # always hop 0: the accumulators start from this partial
                Op, Lp = ops.fwd(q[qi], cur[0][ki], cur[1][ki], True)
                O_acc[qi].copy_(Op); L_acc[qi].copy_(Lp)
            elif o < rank:                                    # all local queries x first half of the visiting block (strided view)
                Op, Lp = ops.fwd(q[qi], cur[0][ki][:, :, :c], cur[1][ki][:, :, :c], False)
                ops.merge_(O_acc[qi], L_acc[qi], Op, Lp, 0)
            else:                                             # second half of the local queries x the whole visiting block
                Op, Lp = ops.fwd(q[qi][:, :, c:], cur[0][ki], cur[1][ki], False)
                ops.merge_(O_acc[qi], L_acc[qi], Op, Lp, c)
        _mark(timeline, q, f"fwd{s}:compute_end")
        if pending is not None:
            pending.wait()
            cur = nxt
        _mark(timeline, q, f"fwd{s}:end")
    return O_acc.to(q.dtype), L_acc


def ring_attention_backward(q, k, v, O, dO, LSE, group=None, ops=None, comm=None, splits=None, timeline=None):
    """Gradients for ring_attention_forward.  Returns (dq, dk, dv) for the local rows, in q.dtype.

    Two rings run under the compute: the K/V block of the next hop is prefetched while the current hop's
    kernels run, and the fp32 dK/dV accumulators of the block that just left are in flight to the next rank
    while this rank already computes its contribution to the following block; they are added on arrival.
    All ring buffers are allocated once (two K/V receive buffers, two accumulator pairs)."""
    ops = ops or CudaOps()
    comm = comm or DistComm(group)
    world, rank = comm.world, comm.rank
    B, H, S2, D = q.shape
    Hk = k.shape[1]; G = H // Hk
    c = S2 // 2
    parts = _parts(B, Hk, G, splits if splits is not None else (4 if world > 1 else 1))
    f32 = dict(dtype=torch.float32, device=q.device)
    q, O, dO, LSE = q.contiguous(), O.contiguous(), dO.contiguous(), LSE.contiguous()
    delta = ops.delta(O, dO)                                  # global: uses the final O of the local rows
    L_hi, d_hi = LSE[:, :, c:].contiguous(), delta[:, :, c:].contiguous()   # the C ABI wants contiguous fp32 statistics
    dq_acc = torch.zeros(B, H, S2, D, **f32)
    cur = [k.contiguous(), v.contiguous()]
    bufs = [[torch.empty_like(cur[0]), torch.empty_like(cur[1])] for _ in range(min(2, world - 1))]
    acc = [torch.zeros(B, Hk, S2, D, **f32), torch.zeros(B, Hk, S2, D, **f32)]   # accumulators of the visiting block
    acc_in = [torch.empty_like(acc[0]), torch.empty_like(acc[1])] if world > 1 else None
    acc_pending = None
    for s in range(world):
        _mark(timeline, q, f"bwd{s}:start")
        kv_pending, nxt = None, None
        if s + 1 < world:
            nxt = bufs[s % 2]
            kv_pending = comm.exchange(cur, nxt)
        o = (rank - s) % world
        grads = []
        for qi, ki in parts:
            if o == rank:
                dq, dk, dv = ops.bwd(q[qi], cur[0][ki], cur[1][ki], O[qi], dO[qi], LSE[qi], delta[qi], True)
                dq_acc[qi].add_(dq); rows = slice(0, S2)
            elif o < rank:
                dq, dk, dv = ops.bwd(q[qi], cur[0][ki][:, :, :c], cur[1][ki][:, :, :c], O[qi], dO[qi], LSE[qi], delta[qi], False)
                dq_acc[qi].add_(dq); rows = slice(0, c)
            else:
                hi = lambda t: t[qi][:, :, c:]
                dq, dk, dv = ops.bwd(hi(q), cur[0][ki], cur[1][ki], hi(O), hi(dO), L_hi[qi], d_hi[qi], False)
                dq_acc[qi][:, :, c:].add_(dq); rows = slice(0, S2)
            grads.append((ki, rows, dk, dv))
        _mark(timeline, q, f"bwd{s}:compute_end")
        if acc_pending is not None:                           # accumulators of this block, sent by the previous rank last hop
            acc_pending.wait()
            acc, acc_in = acc_in, acc
        for ki, rows, dk, dv in grads:
            acc[0][ki][:, :, rows].add_(dk); acc[1][ki][:, :, rows].add_(dv)
        if world > 1:                                         # pass them on with their block (the last hop brings them home)
            acc_pending = comm.exchange(acc, acc_in)
        if kv_pending is not None:
            kv_pending.wait()
            cur = nxt
        _mark(timeline, q, f"bwd{s}:end")
    if acc_pending is not None:
        acc_pending.wait()
        acc = acc_in
    return dq_acc.to(q.dtype), acc[0].to(q.dtype), acc[1].to(q.dtype)


class RingFlashAttentionFunction(torch.autograd.Function):
    """Autograd wrapper with the same saved-tensor contract as FlashAttentionFunction (Q,K,V,O,LSE)."""

    @staticmethod
    def forward(ctx, q, k, v, group=None, comm=None, splits=None, timeline=None):
        assert q.dtype in (torch.float16, torch.bfloat16) and q.ndim == 4
        q_, k_, v_ = q.contiguous(), k.contiguous(), v.contiguous()
        O, LSE = ring_attention_forward(q_, k_, v_, group, None, comm, splits, timeline)
        ctx.save_for_backward(q_, k_, v_, O, LSE)
        ctx.ring = (group, comm, splits, timeline)
        return O

    @staticmethod
    def backward(ctx, dO):
        q, k, v, O, LSE = ctx.saved_tensors
        group, comm, splits, timeline = ctx.ring
        dq, dk, dv = ring_attention_backward(q, k, v, O, dO.contiguous(), LSE, group, None, comm, splits, timeline)
        return dq, dk, dv, None, None, None, None


def ring_flash_attention(q, k, v, group=None, comm=None, splits=None, timeline=None):
    """Causal attention over a sequence sharded zigzag-wise across the ranks of `group` (or of an injected `comm`)."""
    return RingFlashAttentionFunction.apply(q, k, v, group, comm, splits, timeline)


class ThreadRingComm:
    """In-process ring for P simulated ranks, one Python thread each, all on ONE device and one CUDA stream: the ring schedule
    and the real local kernels can be verified on a single GPU (tests/) or on CPU tensors.  exchange() snapshots the send
    buffers (a buffered send) and hands them to the next rank's queue; wait() copies the previous rank's snapshot into the
    receive buffers.  Stream order makes the data dependencies safe: a snapshot is enqueued after its producers and the copy
    after the snapshot, all on the same stream."""

    def __init__(self, rank, world, queues):
        self.rank, self.world, self.queues = rank, world, queues

    @staticmethod
    def make(world):
        import queue
        qs = [queue.Queue() for _ in range(world)]            # qs[r]: messages for rank r (from r - 1)
        return [ThreadRingComm(r, world, qs) for r in range(world)]

    def exchange(self, send, recv):
        self.queues[(self.rank + 1) % self.world].put([t.clone() for t in send])
        comm = self

        class _P:
            def wait(self_inner):
                got = comm.queues[comm.rank].get(timeout=60)
                for dst, src in zip(recv, got):
                    dst.copy_(src)
        return _P()


def run_virtual_ring(world, fn):
    """Run fn(rank, comm) on `world` threads sharing a ThreadRingComm ring; returns the list of results (exceptions re-raised)."""
    import threading
    comms = ThreadRingComm.make(world)
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            out[r] = fn(r, comms[r])
        except BaseException as e:   # noqa: BLE001 - re-raised in the caller
            err[r] = e
    ths = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out
