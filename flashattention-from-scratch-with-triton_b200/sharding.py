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


class _shared_sms:
    """Context: while a sequence-parallel pass runs with the CUDA kernels, transfers overlap the attention launches."""

    def __init__(self, on):
        import os
        self.on = on and os.environ.get("FA_CP_SHARED_SMS", "0") not in ("", "0")      # opt-in: no gain measured so far

    def __enter__(self):
        if self.on:
            from .interface import set_shared_sms
            self.prev = set_shared_sms(True)

    def __exit__(self, *a):
        if self.on:
            from .interface import set_shared_sms
            set_shared_sms(self.prev)


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
    """See _ring_attention_forward_impl; with the CUDA kernels the launches are marked as sharing the GPU with the transfers."""
    with _shared_sms(ops is None and q.is_cuda):
        return _ring_attention_forward_impl(q, k, v, group, ops, comm, splits, timeline)


def _ring_attention_forward_impl(q, k, v, group=None, ops=None, comm=None, splits=None, timeline=None):
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
    """See _ring_attention_backward_impl; with the CUDA kernels the launches are marked as sharing the GPU with the transfers."""
    with _shared_sms(ops is None and q.is_cuda):
        return _ring_attention_backward_impl(q, k, v, O, dO, LSE, group, ops, comm, splits, timeline)


def _ring_attention_backward_impl(q, k, v, O, dO, LSE, group=None, ops=None, comm=None, splits=None, timeline=None):
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


# ------------------------------------------------------------------------------------------------
# NVSwitch variant: gather K/V once, one range-masked launch per head group, scatter-add dK/dV
# ------------------------------------------------------------------------------------------------
# The ring above is the NVLink-ring schedule (P hops, send/recv to the neighbours).  On an NVSwitch box every GPU reaches every peer
# at full bandwidth and 180 GB of HBM hold the whole K/V of a long sequence many times over (C5: 2 x 1 GiB), so the hop structure
# buys nothing: K and V are all-gathered ONCE (per head group, so the transfer of group g+1 runs under the kernels of group g),
# put in global order, and each rank runs the single-GPU kernels ONCE per group on "local queries x all keys" with a Ranges mask
# (query at global position p sees keys [0, p]): no per-hop launches, no (O, LSE) merges, no fp32 accumulators on the wire, and the
# dynamic tile scheduler balances the causal work inside the launch.  The backward's dK/dV partials (16-bit, what the kernel
# writes) go to their owner ranks with one all-to-all per group and are summed there in fp32.


def zigzag_ranges(rank: int, world: int, c: int, B: int, device) -> "object":
    """Ranges for the zigzag-local query rows (chunks `rank` and 2P-1-rank, c rows each) against the GLOBALLY ordered keys."""
    from .interface import Ranges
    a, b = zigzag_chunks(rank, world)
    i = torch.arange(c, device=device)
    gpos = torch.cat([a * c + i, b * c + i])                       # global position of local row
    row_hi = (gpos + 1)[None].expand(B, 2 * c)
    row_lo = torch.zeros_like(row_hi)
    j = torch.arange(2 * c * world, device=device)
    col_lo = torch.where(j < a * c, 0, torch.where(j < (a + 1) * c, j - a * c, torch.where(j < b * c, c, torch.where(
        j < (b + 1) * c, c + j - b * c, 2 * c))))[None].expand(B, -1)
    col_hi = torch.full_like(col_lo, 2 * c)
    return Ranges(row_lo, row_hi, col_lo, col_hi)


def _to_global(buf: torch.Tensor) -> torch.Tensor:
    """[P, B, h, 2c, D] (rank-major, each rank's zigzag pair) -> [B, h, 2Pc, D] in global sequence order."""
    P, B, h, S2, D = buf.shape
    c = S2 // 2
    out = torch.empty(B, h, 2 * P, c, D, dtype=buf.dtype, device=buf.device)
    b5 = buf.view(P, B, h, 2, c, D)
    out[:, :, :P] = b5[:, :, :, 0].permute(1, 2, 0, 3, 4)
    out[:, :, P:] = b5[:, :, :, 1].flip(0).permute(1, 2, 0, 3, 4)
    return out.view(B, h, 2 * P * c, D)


def _from_global(t: torch.Tensor, P: int) -> torch.Tensor:
    """Inverse of _to_global: [B, h, 2Pc, D] -> [P, B, h, 2c, D]."""
    B, h, N, D = t.shape
    c = N // (2 * P)
    t5 = t.view(B, h, 2 * P, c, D)
    out = torch.empty(P, B, h, 2, c, D, dtype=t.dtype, device=t.device)
    out[:, :, :, 0] = t5[:, :, :P].permute(2, 0, 1, 3, 4)
    out[:, :, :, 1] = t5[:, :, P:].flip(2).permute(2, 0, 1, 3, 4)
    return out.view(P, B, h, 2 * c, D)


class GatherOps:
    """Local kernels of the gather variant = the sm_100a library with a Ranges mask."""

    def fwd(self, q, k, v, ranges):
        from .interface import flash_attention_forward
        return flash_attention_forward(q, k, v, False, None, ranges)

    def bwd(self, q, k, v, o, do, lse, ranges):
        from .interface import flash_attention_backward
        return flash_attention_backward(q, k, v, o, do, lse, False, None, ranges)


class DistCollectives:
    """all_gather / all_to_all over torch.distributed (NCCL over NVSwitch on GPUs, gloo on CPU ranks), asynchronous."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group); self.rank = dist.get_rank(group)

    def begin(self, phase, arena_need=0):
        pass

    def all_gather(self, t):
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if self.world == 1:
            out[0].copy_(t); return out, None
        return out, self.dist.all_gather_into_tensor(out.view(-1), t.reshape(-1), group=self.group, async_op=True)

    def all_to_all(self, send):
        recv = torch.empty_like(send)
        if self.world == 1:
            recv.copy_(send); return recv, None
        if send.is_cuda:
            return recv, self.dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group, async_op=True)
        ins = list(send.unbind(0)); outs = list(recv.unbind(0))            # gloo has no all_to_all_single
        self.dist.all_to_all(outs, ins, group=self.group) if self.dist.get_backend(self.group) != "gloo" else _gloo_all_to_all(self.dist, outs, ins, self.group)
        return recv, None


class PeerCollectives:
    """all_gather / all_to_all over NVLink PEER MEMORY, moved by the COPY ENGINES: no SM is taken from the persistent attention
    kernels (an NCCL kernel next to a running attention launch costs it 12-16 % — its CTAs cannot co-reside with CTAs that own the
    whole register file — a copy-engine transfer 3 %; profiles/r02_probe_peer_copy_vs_nccl.json).

    torch symmetric memory provides the peer mappings; the protocol needs no kernel at all:
      writer  waits (cuStreamWaitValue32 on its own flag words) until every reader has acknowledged the slot's previous use, writes
              its data into its slot of the symmetric arena, then stores the epoch into the `ready` word it owns on every peer
              (cuStreamWriteValue32 through the peer mapping) — all on the compute stream, ordered behind the producing kernels;
      reader  on one of a few side streams: waits for `ready[src] >= epoch` in its OWN memory, copies the peer's slot with
              cudaMemcpyAsync (device-to-device peer copy = copy engine), then stores the epoch into its `ack` word on the source.
    Every rank issues the same sequence of collectives, so slots are numbered by call order and `begin(phase)` rewinds the count:
    the same call site gets the same slot (its own symmetric buffer, and flag index) every step."""

    N_SLOTS = 64

    def __init__(self, group=None, n_streams=4):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        from cuda.bindings import driver as drv
        self.drv = drv
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group); self.rank = dist.get_rank(self.group)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.symm, self.arena, self.hdl, self.peer_arena = symm, {}, {}, {}
        nflag = 2 * 2 * self.world * self.N_SLOTS                 # [phase][ready | ack][src / dst rank][slot]
        self.flags = symm.empty(nflag, dtype=torch.int32, device=self.dev)
        self.flags.zero_()
        self.fh = symm.rendezvous(self.flags, self.group.group_name)
        self.peer_flags = [self.fh.get_buffer(r, (nflag,), torch.int32) for r in range(self.world)]
        torch.cuda.synchronize(); dist.barrier(self.group)
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(n_streams)]
        self.epoch = 0; self.phase = None; self.off = 0; self.call = 0
        self.last_use = {}

    def _flag(self, phase, kind, r, slot):                        # index into the int32 flag array
        return (((0 if phase == "fwd" else 1) * 2 + kind) * self.world + r) * self.N_SLOTS + slot

    def begin(self, phase, arena_need=0):
        """Start of a forward / backward pass: rewind the slot counter of that phase."""
        self.epoch += 1; self.phase = phase; self.call = 0

    def _wait_geq(self, stream, flag_index, value):
        addr = self.flags.data_ptr() + 4 * flag_index
        err, = self.drv.cuStreamWaitValue32(stream.cuda_stream, addr, value, self.drv.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ)
        assert err == self.drv.CUresult.CUDA_SUCCESS, err

    def _write(self, stream, peer, flag_index, value):
        addr = self.peer_flags[peer].data_ptr() + 4 * flag_index
        err, = self.drv.cuStreamWriteValue32(stream.cuda_stream, addr, value, 0)
        assert err == self.drv.CUresult.CUDA_SUCCESS, err

    def _slot(self, nbytes):
        """Next slot of the phase: its own symmetric buffer (allocated and exchanged on first use — a collective step, every rank
        gets here in the same order; one buffer per slot keeps every mapping far below 2 GiB)."""
        ph = self.phase
        assert ph is not None, "PeerCollectives.begin(phase) first"
        slot = self.call; self.call += 1
        assert slot < self.N_SLOTS
        key = (ph, slot)
        if key not in self.arena or self.arena[key].numel() < nbytes:
            torch.cuda.synchronize()                                         # nobody may still read a buffer that is replaced
            t = self.symm.empty(nbytes, dtype=torch.uint8, device=self.dev)
            h = self.symm.rendezvous(t, self.group.group_name)
            self.arena[key], self.hdl[key] = t, h
            self.peer_arena[key] = [h.get_buffer(r, (nbytes,), torch.uint8) for r in range(self.world)]
        return ph, slot, 0

    def _publish(self, ph, slot, fill):
        """Writer side: wait for the acks of the slot's previous use, fill the slot, raise `ready` on every peer."""
        cur = torch.cuda.current_stream()
        prev = self.last_use.get((ph, slot), 0)
        if prev:
            for r in range(self.world):
                if r != self.rank:
                    self._wait_geq(cur, self._flag(ph, 1, r, slot), prev)
        fill()
        for r in range(self.world):
            if r != self.rank:
                self._write(cur, r, self._flag(ph, 0, self.rank, slot), self.epoch)
        self.last_use[(ph, slot)] = self.epoch

    def _pull(self, ph, slot, jobs):
        """Reader side: jobs = [(src rank, dst tensor, byte offset in the peer arena, nbytes)], spread over the side streams."""
        cur = torch.cuda.current_stream()
        ev0 = torch.cuda.Event(); ev0.record(cur)
        used = []
        for i, (r, dst, off, nbytes) in enumerate(jobs):
            st = self.streams[i % len(self.streams)]
            st.wait_event(ev0)
            self._wait_geq(st, self._flag(ph, 0, r, slot), self.epoch)
            with torch.cuda.stream(st):
                dst.view(torch.uint8).view(-1).copy_(self.peer_arena[(ph, slot)][r][off:off + nbytes], non_blocking=True)
            self._write(st, r, self._flag(ph, 1, self.rank, slot), self.epoch)
            if st not in used:
                used.append(st)
        evs = []
        for st in used:
            e = torch.cuda.Event(); e.record(st); evs.append(e)
        return _EventWait(evs)

    def all_gather(self, t):
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        ph, slot, off = self._slot(nbytes)
        mine = self.arena[(ph, slot)][off:off + nbytes]
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        self._publish(ph, slot, lambda: mine.copy_(t.view(torch.uint8).view(-1)))
        out[self.rank].copy_(t)
        jobs = [((self.rank + i) % self.world, None, off, nbytes) for i in range(1, self.world)]
        return out, self._pull(ph, slot, [(r, out[r], o, n) for r, _, o, n in jobs])

    def all_to_all(self, send):
        """send [P, ...]: block p goes to rank p.  (The producer's copy into the arena is the only extra pass.)"""
        send = send.contiguous()
        blk = send[0].numel() * send.element_size()
        ph, slot, off = self._slot(blk * self.world)
        mine = self.arena[(ph, slot)][off:off + blk * self.world]
        recv = torch.empty_like(send)
        self._publish(ph, slot, lambda: mine.copy_(send.view(torch.uint8).view(-1)))
        recv[self.rank].copy_(send[self.rank])
        jobs = [((self.rank + i) % self.world, None) for i in range(1, self.world)]
        return recv, self._pull(ph, slot, [(r, recv[r], off + blk * self.rank, blk) for r, _ in jobs])


class _EventWait:
    def __init__(self, events):
        self.events = events

    def wait(self):
        cur = torch.cuda.current_stream()
        for e in self.events:
            cur.wait_event(e)


def _gloo_all_to_all(dist, outs, ins, group):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ops = []
    for r in range(world):
        if r == rank:
            outs[r].copy_(ins[r])
        else:
            ops.append(dist.P2POp(dist.isend, ins[r].contiguous(), r, group)); ops.append(dist.P2POp(dist.irecv, outs[r], r, group))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()


def _head_groups(Hk: int, G: int, n):
    """(query-head slice, K/V-head slice) per group.  `n` = number of equal groups, or a list of K/V heads per group."""
    if isinstance(n, int):
        n = max(1, min(n, Hk)); per = -(-Hk // n)
        sizes = [min(per, Hk - h) for h in range(0, Hk, per)]
    else:
        sizes = [int(x) for x in n if int(x) > 0]
        assert sum(sizes) == Hk, f"head groups {sizes} do not add up to {Hk} K/V heads"
    out, h = [], 0
    for sz in sizes:
        out.append((slice(h * G, (h + sz) * G), slice(h, h + sz))); h += sz
    return out


def default_head_groups(Hk: int):
    """Small group first, then big ones: the first group's all-gather is the only exposed one in the forward, and the backward
    walks the groups in REVERSE order so that the only exposed all-to-all (the last) is the small group's.  32 heads -> [2, 6, 8, 8, 8]."""
    if Hk < 8:
        return [1] * Hk if Hk <= 2 else [1, Hk - 1]
    a = max(1, Hk // 16); b = max(1, 3 * Hk // 16); rest = Hk - a - b
    q = rest // 3
    return [a, b, q, q, rest - 2 * q]


def make_cp_group(max_ctas: int = 8, ranks=None):
    """A dedicated NCCL communicator for the sequence-parallel collectives whose kernels are capped at `max_ctas` CTAs: the
    attention kernels are persistent (one CTA per SM, whole register file), so every SM an NCCL kernel occupies is an SM the
    running attention launch does not get; the transfers are small next to the compute they hide under."""
    import torch.distributed as dist
    if dist.get_backend() != "nccl":
        return dist.new_group(ranks=ranks)
    opts = dist.ProcessGroupNCCL.Options()
    opts.config.max_ctas = int(max_ctas); opts.config.min_ctas = 1
    return dist.new_group(ranks=ranks, pg_options=opts)


def gather_attention_forward(q, k, v, group=None, ops=None, coll=None, groups=None, timeline=None):
    """See _gather_attention_forward_impl; with the CUDA kernels the launches are marked as sharing the GPU with the transfers."""
    with _shared_sms(ops is None and q.is_cuda):
        return _gather_attention_forward_impl(q, k, v, group, ops, coll, groups, timeline)


def _gather_attention_forward_impl(q, k, v, group=None, ops=None, coll=None, groups=None, timeline=None):
    """Causal attention over the global sequence; q, k, v are this rank's zigzag-local [B,H,2c,D] / [B,Hk,2c,D] tensors.
    Returns (O, LSE, saved) where `saved` carries the gathered K/V (global order, per head group) and the Ranges for the backward."""
    ops = ops or GatherOps()
    coll = coll or DistCollectives(group)
    P, rank = coll.world, coll.rank
    B, H, S2, D = q.shape
    Hk = k.shape[1]; G = H // Hk
    c = S2 // 2
    ranges = zigzag_ranges(rank, P, c, B, q.device)
    hg = _head_groups(Hk, G, groups if groups is not None else default_head_groups(Hk))
    coll.begin("fwd", (k.numel() + v.numel()) * k.element_size() + 512 * len(hg))
    _mark(timeline, q, "fwd:start")
    pend = [(coll.all_gather(k[:, ks].contiguous()), coll.all_gather(v[:, ks].contiguous())) for _, ks in hg]   # all posted up front
    O = torch.empty_like(q); LSE = torch.empty(B, H, S2, dtype=torch.float32, device=q.device)
    kv_glob = []
    for gi, (qs, ks) in enumerate(hg):
        (kb, kw), (vb, vw) = pend[gi]
        for w in (kw, vw):
            if w is not None:
                w.wait()
        Kg, Vg = _to_global(kb), _to_global(vb)
        pend[gi] = None
        Og, Lg = ops.fwd(q[:, qs], Kg, Vg, ranges)
        O[:, qs] = Og; LSE[:, qs] = Lg
        kv_glob.append((Kg, Vg))
        _mark(timeline, q, f"fwd:group{gi}")
    return O, LSE, (kv_glob, ranges, hg)


def gather_attention_backward(q, O, dO, LSE, saved, group=None, ops=None, coll=None, timeline=None):
    """See _gather_attention_backward_impl; with the CUDA kernels the launches are marked as sharing the GPU with the transfers."""
    with _shared_sms(ops is None and q.is_cuda):
        return _gather_attention_backward_impl(q, O, dO, LSE, saved, group, ops, coll, timeline)


def _gather_attention_backward_impl(q, O, dO, LSE, saved, group=None, ops=None, coll=None, timeline=None):
    """Gradients for gather_attention_forward.  dQ is local; the dK / dV partials of every key (16-bit, global order) are sent to
    the key's owner (one all-to-all per head group, posted as soon as the group's kernels are enqueued) and summed there in fp32."""
    ops = ops or GatherOps()
    coll = coll or DistCollectives(group)
    P = coll.world
    kv_glob, ranges, hg = saved
    dq = torch.empty_like(q)
    Hk = sum(ks.stop - ks.start for _, ks in hg)
    B, H, S2, D = q.shape
    dk = torch.empty(B, Hk, S2, D, dtype=q.dtype, device=q.device); dv = torch.empty_like(dk)
    coll.begin("bwd", 2 * P * dk.numel() * dk.element_size() + 512 * len(hg))
    _mark(timeline, q, "bwd:start")
    pend = {}
    order = list(range(len(hg)))[::-1]                        # reverse: the last (exposed) all-to-all belongs to the smallest group

    def finish(gi):                                           # partials of group gi have arrived: sum them in fp32, round once
        (kb, kw), (vb, vw) = pend.pop(gi)
        for w in (kw, vw):
            if w is not None:
                w.wait()
        ks = hg[gi][1]
        dk[:, ks] = torch.sum(kb, dim=0, dtype=torch.float32).to(q.dtype); dv[:, ks] = torch.sum(vb, dim=0, dtype=torch.float32).to(q.dtype)

    for n, gi in enumerate(order):
        qs, ks = hg[gi]
        Kg, Vg = kv_glob[gi]
        dqg, dKg, dVg = ops.bwd(q[:, qs], Kg, Vg, O[:, qs], dO[:, qs], LSE[:, qs], ranges)
        dq[:, qs] = dqg
        pend[gi] = (coll.all_to_all(_from_global(dKg, P)), coll.all_to_all(_from_global(dVg, P)))
        if n >= 1:
            finish(order[n - 1])                              # the previous group's transfer ran under this group's kernels
        _mark(timeline, q, f"bwd:group{gi}")
    finish(order[-1])
    _mark(timeline, q, "bwd:end")
    return dq, dk, dv


class GatherFlashAttentionFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, group=None, coll=None, groups=None, timeline=None):
        assert q.dtype in (torch.float16, torch.bfloat16) and q.ndim == 4
        q_ = q.contiguous()
        O, LSE, saved = gather_attention_forward(q_, k, v, group, None, coll, groups, timeline)
        ctx.save_for_backward(q_, O, LSE)
        ctx.cp = (saved, group, coll, timeline)
        return O

    @staticmethod
    def backward(ctx, dO):
        q, O, LSE = ctx.saved_tensors
        saved, group, coll, timeline = ctx.cp
        dq, dk, dv = gather_attention_backward(q, O, dO.contiguous(), LSE, saved, group, None, coll, timeline)
        return dq, dk, dv, None, None, None, None


def gather_flash_attention(q, k, v, group=None, coll=None, groups=None, timeline=None):
    """Causal attention over a zigzag-sharded sequence, NVSwitch variant (see the section comment): same inputs / outputs as
    ring_flash_attention."""
    return GatherFlashAttentionFunction.apply(q, k, v, group, coll, groups, timeline)


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


class ThreadCollectives:
    """In-process all_gather / all_to_all for P simulated ranks (threads of one process, one device, one CUDA stream): a
    thread barrier orders the enqueues, the stream orders the data."""

    def __init__(self, rank, world, shared):
        self.rank, self.world, self.shared = rank, world, shared

    @staticmethod
    def make(world):
        import threading
        shared = dict(barrier=threading.Barrier(world, timeout=120), slots=[None] * world)
        return [ThreadCollectives(r, world, shared) for r in range(world)]

    def begin(self, phase, arena_need=0):
        pass

    def _exchange(self, t, pick):
        sh_ = self.shared
        sh_["slots"][self.rank] = t
        sh_["barrier"].wait()
        out = pick(sh_["slots"])
        sh_["barrier"].wait()                                # nobody overwrites a slot that is still being read
        return out, None

    def all_gather(self, t):
        return self._exchange(t, lambda slots: torch.stack(list(slots)))

    def all_to_all(self, send):
        return self._exchange(send, lambda slots: torch.stack([slots[r][self.rank] for r in range(self.world)]))


def run_virtual_ring(world, fn, make=None):
    """Run fn(rank, comm) on `world` threads sharing a ThreadRingComm ring (or the comm objects of `make(world)`); returns the list
    of results (exceptions re-raised)."""
    import threading
    comms = (make or ThreadRingComm.make)(world)
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
