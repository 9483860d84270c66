"""Tests that need >= 2 GPUs on one box: `gpurun --gpus 2 -- python -m pytest tests -m multigpu`.
Batch x head sharding and the zigzag K/V ring over NCCL P2P with the sm_100a kernels as the local op; a second device in one
process.  Marked `multigpu`, NOT `gpu`: the driver's 1-GPU `-m gpu` run selects none of them (nothing to skip there); the
ring's CUDA path is covered on one GPU by tests/test_gpu_multi.py::test_virtual_ring_on_one_gpu."""
import os
import socket

import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.multigpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _close(x, ref, atol=1e-2, rtol=1e-2):
    return torch.allclose(x.float(), ref.float(), atol=atol, rtol=rtol)


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    import flashattn_b200 as fa
    import flashattn_b200.sharding as sh
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        out = {}
        # ---- batch x head sharding: local slice == slice of the single-GPU result, no collective
        Q, K, V, dO = orc.make_inputs(2, 4, 512, 512, 128, torch.bfloat16, seed=5)
        O_loc = sh.sharded_flash_attention(Q.cuda(), K.cuda(), V.cuda(), True, rank, world)
        O_full = fa.flash_attention(Q.cuda(), K.cuda(), V.cuda(), True)
        out["shard_bitwise"] = bool(torch.equal(O_loc, sh.local_shard(O_full, rank, world)))
        # ---- zigzag ring, causal, forward + backward through autograd
        S = 512 * world
        Q, K, V, dO = orc.make_inputs(1, 2, S, S, 128, torch.bfloat16, seed=6)
        q, k, v, do = (sh.zigzag_split(t, rank, world).cuda() for t in (Q, K, V, dO))
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        O = sh.ring_flash_attention(q, k, v)
        O.backward(do)
        torch.cuda.synchronize()
        rO, rLSE, rdQ, rdK, rdV = orc.closed_form(Q, K, V, dO, True, dtype=torch.float32)
        errs = {}
        for name, x, r in (("O", O, rO), ("dQ", q.grad, rdQ), ("dK", k.grad, rdK), ("dV", v.grad, rdV)):
            r = sh.zigzag_split(r, rank, world)
            d = (x.detach().cpu().float() - r).abs()
            errs[name] = float((d / (1e-2 + 1e-2 * r.abs())).max())
        out["ring_norm_err"] = errs
        # the one-shot kernel on the whole sequence, same inputs: the yardstick for the sharded variants' error
        qf, kf, vf = (t.cuda().requires_grad_(True) for t in (Q, K, V))
        Of = fa.flash_attention(qf, kf, vf, True); Of.backward(dO.cuda())
        out["one_shot_norm_err"] = {n: float(((x.detach().cpu().float() - r).abs() / (1e-2 + 1e-2 * r.abs())).max())
                                    for n, x, r in (("O", Of, rO), ("dQ", qf.grad, rdQ), ("dK", kf.grad, rdK), ("dV", vf.grad, rdV))}
        # ---- NVSwitch variant (gather): NCCL collectives, then peer memory + copy engines; two steps each (slot reuse, acks)
        for name, coll in (("gather_nccl", None), ("gather_peer", sh.PeerCollectives())):
            for step in range(2):
                q.grad = None; k.grad = None; v.grad = None
                O = sh.gather_flash_attention(q, k, v, None, coll, [1, 1])
                O.backward(do)
            torch.cuda.synchronize()
            errs = {}
            for tn, x, r in (("O", O, rO), ("dQ", q.grad, rdQ), ("dK", k.grad, rdK), ("dV", v.grad, rdV)):
                r = sh.zigzag_split(r, rank, world)
                d = (x.detach().cpu().float() - r).abs()
                errs[tn] = float((d / (1e-2 + 1e-2 * r.abs())).max())
            out[name + "_norm_err"] = errs
        ret[rank] = out
    finally:
        dist.destroy_process_group()


def test_sharding_and_ring_on_gpus():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        for r in range(world):
            assert ret[r]["shard_bitwise"], r
            # sharded results: inside the contract (atol = rtol = 1e-2) or no worse than 1.1 x the one-shot kernel on the same inputs
            bound = max(1.0, 1.1 * max(ret[r]["one_shot_norm_err"].values()))
            for variant in ("ring", "gather_nccl", "gather_peer"):
                assert max(ret[r][variant + "_norm_err"].values()) <= bound, (r, variant, ret[r])


def test_second_device_in_the_same_process():
    """The > 48 KB dynamic shared memory opt-in is a per-device kernel attribute: the first launch on cuda:1 of a process that
    has already used cuda:0 must work (fwd + both backward structures)."""
    import flashattn_b200 as fa
    for dev in ("cuda:0", "cuda:1"):
        for D in (64, 128):
            Q, K, V, dO = (t.to(dev) for t in orc.make_inputs(1, 2, 256, 256, D, torch.bfloat16, seed=1))
            q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
            O = fa.flash_attention(q, k, v, True); O.backward(dO)
            rO, _, rdQ, rdK, rdV = orc.closed_form(Q.cpu(), K.cpu(), V.cpu(), dO.cpu(), True)
            for x, ref in ((O, rO), (q.grad, rdQ), (k.grad, rdK), (v.grad, rdV)):
                assert _close(x.detach().cpu(), ref), (dev, D)
