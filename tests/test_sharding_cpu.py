"""Host-side multi-GPU logic on CPU ranks (gloo, world_size 2 and 4): batch x head partitioning and the
zigzag ring schedule (forward merge + backward with travelling dK/dV), with the CPU oracle injected as
the local kernel.  The CUDA library is not involved here; test_gpu_multi.py covers it on real GPUs."""
import math
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import flashattn_b200.sharding as sh
from oracle import attention_oracle as orc


class OracleOps:
    """Local 'kernels' for CPU ranks: exact fp32 math with the same call contract as sharding.CudaOps."""

    def fwd(self, q, k, v, causal):
        O, LSE = orc.closed_form(q, k, v, None, causal, dtype=torch.float32)
        return O.to(q.dtype), LSE

    def merge_(self, O_acc, LSE_acc, O_part, LSE_part, q_off):
        n = O_part.shape[2]
        O, L = orc.merge_partials(O_acc[:, :, q_off:q_off + n], LSE_acc[:, :, q_off:q_off + n], O_part, LSE_part)
        O_acc[:, :, q_off:q_off + n] = O; LSE_acc[:, :, q_off:q_off + n] = L

    def delta(self, O, dO):
        return (O.float() * dO.float()).sum(-1)

    def bwd(self, q, k, v, o, do, lse, delta, causal):
        D = q.shape[-1]; scale = 1 / math.sqrt(D)
        qf, kf, vf, dof = q.float(), k.float(), v.float(), do.float()
        S = qf @ kf.transpose(-1, -2) * scale
        if causal:
            i = torch.arange(q.shape[2]); j = torch.arange(k.shape[2])
            S = S.masked_fill(~(i[:, None] >= j[None, :]), float("-inf"))
        P = torch.exp(S - lse[..., None])
        dV = P.transpose(-1, -2) @ dof
        dS = P * (dof @ vf.transpose(-1, -2) - delta[..., None])
        return (dS @ kf * scale).to(q.dtype), (dS.transpose(-1, -2) @ qf * scale).to(q.dtype), dV.to(q.dtype)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _ring_worker(rank, world, port, S, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        Q, K, V, dO = orc.make_inputs(1, 2, S, S, 64, torch.float32, seed=3)
        q, k, v, do = (sh.zigzag_split(t, rank, world) for t in (Q, K, V, dO))
        ops = OracleOps()
        O, LSE = sh.ring_attention_forward(q, k, v, None, ops)
        dq, dk, dv = sh.ring_attention_backward(q, k, v, O, do, LSE, None, ops)
        rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, True, dtype=torch.float64)
        errs = [(a - sh.zigzag_split(b.float(), rank, world)).abs().max().item()
                for a, b in ((O, rO), (dq, rdQ), (dk, rdK), (dv, rdV))]
        errs.append((LSE - sh.zigzag_split(rLSE.float(), rank, world)).abs().max().item())
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_zigzag_ring_fwd_bwd_equals_full_causal_attention(world):
    S = 64 * world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_ring_worker, args=(world, _free_port(), S, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert max(ret[r]) < 2e-5, (r, ret[r])


@pytest.mark.parametrize("world,splits", [(2, 1), (4, 2), (8, 3)])
def test_virtual_ring_threads_equals_full_causal_attention(world, splits):
    """The same ring code driven by ThreadRingComm (P simulated ranks as threads of ONE process — the harness the GPU test
    uses with the CUDA kernels), with sub-launch splitting, against the fp64 closed form."""
    S = 32 * world
    Q, K, V, dO = orc.make_inputs(2 if world == 4 else 1, 3, S, S, 64, torch.float32, seed=9)
    rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, True, dtype=torch.float64)

    def rank_fn(rank, comm):
        q, k, v, do = (sh.zigzag_split(t, rank, world) for t in (Q, K, V, dO))
        ops = OracleOps()
        O, LSE = sh.ring_attention_forward(q, k, v, None, ops, comm, splits)
        return (O, LSE) + tuple(sh.ring_attention_backward(q, k, v, O, do, LSE, None, ops, comm, splits))

    outs = sh.run_virtual_ring(world, rank_fn)
    for i, ref in enumerate((rO, rLSE, rdQ, rdK, rdV)):
        full = sh.zigzag_merge([o[i] for o in outs], dim=2)
        assert (full - ref.float()).abs().max() < 2e-5, i


class OracleGatherOps:
    """Local 'kernels' of the gather variant on CPU ranks: exact fp32 attention under a Ranges row mask (same contract as
    sharding.GatherOps)."""

    @staticmethod
    def _mask(ranges, Sk):
        j = torch.arange(Sk)[None, None, :]
        return (j >= ranges.row_lo[:, :, None]) & (j < ranges.row_hi[:, :, None])          # [B, Sq, Sk]

    def fwd(self, q, k, v, ranges):
        scale = 1 / math.sqrt(q.shape[-1])
        S = (q.float() @ k.float().transpose(-1, -2) * scale).masked_fill(~self._mask(ranges, k.shape[2])[:, None], float("-inf"))
        LSE = torch.logsumexp(S, -1)
        return (torch.exp(S - LSE[..., None]) @ v.float()).to(q.dtype), LSE

    def bwd(self, q, k, v, o, do, lse, ranges):
        scale = 1 / math.sqrt(q.shape[-1])
        qf, kf, vf, dof = q.float(), k.float(), v.float(), do.float()
        S = (qf @ kf.transpose(-1, -2) * scale).masked_fill(~self._mask(ranges, k.shape[2])[:, None], float("-inf"))
        P = torch.exp(S - lse[..., None])
        delta = (o.float() * dof).sum(-1)
        dS = P * (dof @ vf.transpose(-1, -2) - delta[..., None])
        return (dS @ kf * scale).to(q.dtype), (dS.transpose(-1, -2) @ qf * scale).to(q.dtype), (P.transpose(-1, -2) @ dof).to(q.dtype)


def _gather_worker(rank, world, port, S, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        Q, K, V, dO = orc.make_inputs(1, 4, S, S, 64, torch.float32, seed=4)
        q, k, v, do = (sh.zigzag_split(t, rank, world) for t in (Q, K, V, dO))
        sh.zigzag_ranges(rank, world, S // (2 * world), 1, "cpu").validate()               # row side == key side of the mask
        ops = OracleGatherOps()
        O, LSE, saved = sh.gather_attention_forward(q, k, v, None, ops, None, 2)
        dq, dk, dv = sh.gather_attention_backward(q, O, do, LSE, saved, None, ops)
        rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, True, dtype=torch.float64)
        errs = [(a - sh.zigzag_split(b.float(), rank, world)).abs().max().item()
                for a, b in ((O, rO), (dq, rdQ), (dk, rdK), (dv, rdV))]
        errs.append((LSE - sh.zigzag_split(rLSE.float(), rank, world)).abs().max().item())
        ret[rank] = errs
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_gather_variant_fwd_bwd_equals_full_causal_attention(world):
    """NVSwitch variant (all-gather K/V, one range-masked launch per head group, all-to-all + fp32 sum of dK/dV) on gloo ranks."""
    S = 64 * world
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gather_worker, args=(world, _free_port(), S, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert max(ret[r]) < 2e-5, (r, ret[r])


def test_gather_variant_threads_equals_full_causal_attention():
    """The same code driven by ThreadCollectives (8 simulated ranks as threads of one process)."""
    world, S = 8, 256
    Q, K, V, dO = orc.make_inputs(2, 3, S, S, 64, torch.float32, seed=10)
    rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, True, dtype=torch.float64)

    def rank_fn(rank, coll):
        q, k, v, do = (sh.zigzag_split(t, rank, world) for t in (Q, K, V, dO))
        ops = OracleGatherOps()
        O, LSE, saved = sh.gather_attention_forward(q, k, v, None, ops, coll, 3)
        return (O, LSE) + tuple(sh.gather_attention_backward(q, O, do, LSE, saved, None, ops, coll))

    outs = sh.run_virtual_ring(world, rank_fn, sh.ThreadCollectives.make)
    for i, ref in enumerate((rO, rLSE, rdQ, rdK, rdV)):
        full = sh.zigzag_merge([o[i] for o in outs], dim=2)
        assert (full - ref.float()).abs().max() < 2e-5, i


def _shard_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, K, V = orc.make_inputs(2, 3, 64, 64, 64, torch.float32, seed=1, with_dO=False)
        attn = lambda q, k, v, c: orc.closed_form(q, k, v, None, c, dtype=torch.float32)[0]
        O_loc = sh.sharded_flash_attention(Q, K, V, True, rank, world, attn)
        # data path has no collective; gathering here is only the test's check
        outs = [None] * world
        dist.all_gather_object(outs, O_loc)
        if rank == 0:
            full = torch.cat(outs, dim=1).reshape(2, 3, 64, 64)
            ret["err"] = (full - attn(Q, K, V, True)).abs().max().item()
    finally:
        dist.destroy_process_group()


def test_batch_head_sharding_two_ranks():
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_shard_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        assert ret["err"] < 1e-6


def test_partition_is_balanced_and_complete():
    for B, H, W in ((16, 32, 8), (1, 32, 8), (2, 3, 4), (1, 1, 2)):
        parts = sh.partition_batch_heads(B, H, W)
        assert parts[0][0] == 0 and parts[-1][1] == B * H
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_zigzag_split_roundtrip_and_balance():
    t = torch.arange(2 * 3 * 64 * 4, dtype=torch.float32).reshape(2, 3, 64, 4)
    for W in (2, 4, 8):
        parts = [sh.zigzag_split(t, r, W) for r in range(W)]
        assert torch.equal(sh.zigzag_merge(parts), t)
        # causal work (number of visible (q,k) chunk pairs) is identical on every rank
        work = []
        for r in range(W):
            a, b = sh.zigzag_chunks(r, W)
            work.append((a + 1) + (b + 1))
        assert len(set(work)) == 1


# ---------------------------------------------------------------------------------------------------------------------
# Range masks: the kernels take an item's tile range from the FIRST and LAST row of its tile (csrc/fa_fwd.cuh::fwd_item_iters,
# the dK/dV decode in csrc/fa_bwd.cuh).  That is only right because the ranges are monotone; this restates the two formulas and
# checks by brute force that no visible (query, key) pair ever falls outside the tiles a kernel would visit.
# ---------------------------------------------------------------------------------------------------------------------
def _fwd_item_iters(row_lo, row_hi, q0, t, Sq, Sk, causal):
    jb = int(row_lo[min(q0, Sq - 1)]) >> 7
    r0 = q0 + t * 128
    if r0 >= Sq:
        return jb, 0
    last = min(r0 + 127, Sq - 1)
    hi = min(int(row_hi[last]), Sk)
    if causal:
        hi = min(hi, last + 1)
    return jb, max(((hi - 1) >> 7) - jb + 1, 0)


def _dkv_item_range(col_lo, col_hi, jt, Sq, Sk, causal):
    n_qtiles = (Sq + 127) // 128
    i_start = jt if causal else 0
    i_start = max(i_start, int(col_lo[min(jt * 128, Sk - 1)]) >> 7)
    i_end = min(n_qtiles, ((min(int(col_hi[min(jt * 128 + 127, Sk - 1)]), Sq) - 1) >> 7) + 1)
    return i_start, i_end


@pytest.mark.parametrize("causal", [False, True])
def test_range_mask_tile_ranges_cover_every_visible_pair(causal):
    import flashattn_b200 as fa
    rng = torch.Generator().manual_seed(11)
    cases = []
    for _ in range(6):                                   # random packings: boundaries inside tiles, on tile edges, 1-token sequences
        lens = torch.randint(1, 400, (int(torch.randint(2, 9, (1,), generator=rng)),), generator=rng).tolist()
        cu = [0]
        for n in lens:
            cu.append(cu[-1] + n)
        cases.append((fa.Ranges.from_cu_seqlens(cu, cu[-1]), cu[-1], cu[-1]))
    cases.append((fa.Ranges.from_key_padding([77, 300], 300, 300), 300, 300))
    if causal:
        cases.append((fa.Ranges.sliding_window(1, 700, 130), 700, 700))
    for r, Sq, Sk in cases:
        for b in range(r.row_lo.shape[0]):
            lo, hi, clo, chi = r.row_lo[b], r.row_hi[b], r.col_lo[b], r.col_hi[b]
            i = torch.arange(Sq)[:, None]; j = torch.arange(Sk)[None, :]
            vis = (j >= lo[:, None]) & (j < hi[:, None])
            if causal:
                vis &= (i >= j)
            assert torch.equal(vis, (i >= clo[None, :]) & (i < chi[None, :]) & ((i >= j) if causal else True))
            # forward / dQ orientation: 256-row items, two 128-row tiles sharing one K/V stream that starts at tile jb
            for q0 in range(0, Sq, 256):
                for t in (0, 1):
                    jb, nt = _fwd_item_iters(lo, hi, q0, t, Sq, Sk, causal)
                    rows = vis[q0 + t * 128:q0 + t * 128 + 128]
                    if rows.numel() == 0:
                        assert nt == 0
                        continue
                    keys = torch.nonzero(rows.any(0)).flatten()
                    if keys.numel():
                        assert jb * 128 <= int(keys.min()) and int(keys.max()) < (jb + nt) * 128, (q0, t, jb, nt)
            # dK/dV orientation: one 128-row kv tile, q tiles [i_start, i_end)
            for jt in range((Sk + 127) // 128):
                i_start, i_end = _dkv_item_range(clo, chi, jt, Sq, Sk, causal)
                qs = torch.nonzero(vis[:, jt * 128:jt * 128 + 128].any(1)).flatten()
                if qs.numel():
                    assert i_start * 128 <= int(qs.min()) and int(qs.max()) < i_end * 128, (jt, i_start, i_end)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_zigzag_ranges_and_head_groups(world):
    """Host logic of the gather variant: the Ranges of every rank describe ONE mask from both sides (validate()), together they cover
    the causal mask exactly once, and the head groups partition the K/V heads in order."""
    c = 5
    N = 2 * c * world
    seen = torch.zeros(N, N, dtype=torch.int32)
    for rank in range(world):
        r = sh.zigzag_ranges(rank, world, c, 2, "cpu").validate()
        a, b = sh.zigzag_chunks(rank, world)
        gpos = torch.cat([a * c + torch.arange(c), b * c + torch.arange(c)])
        j = torch.arange(N)
        vis = (j[None, :] >= r.row_lo[0][:, None]) & (j[None, :] < r.row_hi[0][:, None])          # [2c, N]
        seen[gpos] += vis.int()
    causal = (torch.arange(N)[:, None] >= torch.arange(N)[None, :]).int()
    assert torch.equal(seen, causal)
    for Hk in (1, 2, 3, 8, 16, 32, 40):
        sizes = sh.default_head_groups(Hk)
        assert sum(sizes) == Hk and all(s > 0 for s in sizes) and sizes[0] == min(sizes)
        for G in (1, 4):
            hg = sh._head_groups(Hk, G, sizes)
            assert [ks.stop - ks.start for _, ks in hg] == sizes and hg[0][1].start == 0 and hg[-1][1].stop == Hk
            assert all(qs.start == ks.start * G and qs.stop == ks.stop * G for qs, ks in hg)
    assert [ks.stop - ks.start for _, ks in sh._head_groups(10, 1, 4)] == [3, 3, 3, 1]
    t = torch.arange(world * 2 * 3 * 2 * c * 4, dtype=torch.float32).view(world, 2, 3, 2 * c, 4)
    assert torch.equal(sh._from_global(sh._to_global(t), world), t)
