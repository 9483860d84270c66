"""Multi-GPU code paths that run on ONE GPU: the ring schedule with the real kernels for P simulated ranks (threads joined by an
in-process ring) and the host-buffer pipeline.  The tests that need >= 2 devices live in tests/test_multigpu.py (`-m multigpu`)."""
import pytest
import torch

from oracle import attention_oracle as orc

pytestmark = pytest.mark.gpu


def test_host_pipeline_equals_monolithic_call():
    import flashattn_b200 as fa
    Q, K, V, dO = orc.make_inputs(2, 6, 512, 512, 128, torch.bfloat16, seed=2)
    hin = [t.pin_memory() for t in (Q, K, V, dO)]
    O, dQ, dK, dV = fa.flash_attention_host(*hin, is_causal=True, chunks=5)
    q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    Or = fa.flash_attention(q, k, v, True); Or.backward(dO.cuda())
    for a, b in ((O, Or), (dQ, q.grad), (dK, k.grad), (dV, v.grad)):
        assert torch.equal(a, b.detach().cpu())            # bitwise: kernels never mix (b,h) pairs
    # pipeline object is reusable
    pipe = fa.HostAttentionPipeline(2, 6, 512, 512, 128, torch.bfloat16, True, chunks=3)
    outs = [torch.empty_like(t).pin_memory() for t in (Q, Q, K, V)]
    for _ in range(2):
        pipe.run(*hin, *outs); torch.cuda.synchronize()
    assert torch.equal(outs[0], O) and torch.equal(outs[3], dV)


@pytest.mark.parametrize("world,D,Hk", [(2, 128, 4), (4, 64, 4), (8, 128, 2)], ids=["P2_d128", "P4_d64", "P8_d128_gqa"])
def test_virtual_ring_on_one_gpu(world, D, Hk):
    """The ring schedule with the REAL local kernels (sharding.CudaOps -> libfa_sm100.so) for P simulated ranks on ONE GPU:
    the ranks are threads of this process joined by sharding.ThreadRingComm (the same ring_attention_forward / _backward code
    that runs over NCCL; only the neighbour exchange is in-process).  Runs on a 1-GPU box, so the ring's CUDA path is covered by
    every GPU test run.  Checked against the fp64 closed form (contract tolerance) and against the one-shot kernel."""
    import flashattn_b200 as fa
    import flashattn_b200.sharding as sh
    B, H, c = 1, 4, 256
    S = 2 * c * world
    g = torch.Generator().manual_seed(60 + world)
    Q = torch.randn(B, H, S, D, generator=g).bfloat16(); dO = torch.randn(B, H, S, D, generator=g).bfloat16()
    K = torch.randn(B, Hk, S, D, generator=g).bfloat16(); V = torch.randn(B, Hk, S, D, generator=g).bfloat16()

    def rank_fn(rank, comm):
        torch.cuda.set_device(0)
        # ring_attention_forward / _backward called directly: autograd would run every simulated rank's backward on the ONE
        # autograd worker thread of the device, where they cannot wait for each other
        q, k, v, do = (sh.zigzag_split(t, rank, world).cuda() for t in (Q, K, V, dO))
        O, LSE = sh.ring_attention_forward(q, k, v, None, None, comm, 2)        # two sub-launches per hop
        dq, dk, dv = sh.ring_attention_backward(q, k, v, O, do, LSE, None, None, comm, 2)
        return O.cpu(), dq.cpu(), dk.cpu(), dv.cpu()

    outs = sh.run_virtual_ring(world, rank_fn)
    torch.cuda.synchronize()
    G = H // Hk
    rO, _, rdQ, rdKe, rdVe = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, True)
    rdK = rdKe.reshape(B, Hk, G, S, D).sum(2); rdV = rdVe.reshape(B, Hk, G, S, D).sum(2)
    q1, k1, v1 = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    O1 = fa.flash_attention(q1, k1, v1, True); O1.backward(dO.cuda())
    import math
    for i, (name, ref, one) in enumerate((("O", rO, O1), ("dQ", rdQ, q1.grad), ("dK", rdK, k1.grad), ("dV", rdV, v1.grad))):
        full = sh.zigzag_merge([o[i] for o in outs], dim=2).float()
        atol = 1e-2 * (math.sqrt(G) if name in ("dK", "dV") else 1.0)
        assert torch.allclose(full, ref.float(), atol=atol, rtol=1e-2), (name, (full - ref.float()).abs().max().item())
        assert torch.allclose(full, one.detach().cpu().float(), atol=2 * atol, rtol=2e-2), name   # two 16-bit results, each within the contract


@pytest.mark.parametrize("world,D,Hk,groups", [(2, 128, 4, 2), (4, 64, 4, 1), (8, 128, 2, 2)], ids=["P2_d128", "P4_d64", "P8_d128_gqa"])
def test_virtual_gather_variant_on_one_gpu(world, D, Hk, groups):
    """The NVSwitch variant (sharding.gather_attention_forward / _backward: all-gather K/V, ONE range-masked launch per head group,
    all-to-all + fp32 sum of the dK/dV partials) with the REAL kernels for P simulated ranks on one GPU (threads joined by
    sharding.ThreadCollectives).  Against the fp64 closed form (contract tolerance) and the one-shot kernel."""
    import math
    import flashattn_b200 as fa
    import flashattn_b200.sharding as sh
    B, H, c = 1, 4, 256
    S = 2 * c * world
    g = torch.Generator().manual_seed(80 + world)
    Q = torch.randn(B, H, S, D, generator=g).bfloat16(); dO = torch.randn(B, H, S, D, generator=g).bfloat16()
    K = torch.randn(B, Hk, S, D, generator=g).bfloat16(); V = torch.randn(B, Hk, S, D, generator=g).bfloat16()

    def rank_fn(rank, coll):
        torch.cuda.set_device(0)
        q, k, v, do = (sh.zigzag_split(t, rank, world).cuda() for t in (Q, K, V, dO))
        O, LSE, saved = sh.gather_attention_forward(q, k, v, None, None, coll, groups)
        dq, dk, dv = sh.gather_attention_backward(q, O, do, LSE, saved, None, None, coll)
        return O.cpu(), dq.cpu(), dk.cpu(), dv.cpu()

    outs = sh.run_virtual_ring(world, rank_fn, sh.ThreadCollectives.make)
    torch.cuda.synchronize()
    G = H // Hk
    rO, _, rdQ, rdKe, rdVe = orc.closed_form(Q, K.repeat_interleave(G, dim=1), V.repeat_interleave(G, dim=1), dO, True)
    rdK = rdKe.reshape(B, Hk, G, S, D).sum(2); rdV = rdVe.reshape(B, Hk, G, S, D).sum(2)
    q1, k1, v1 = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    O1 = fa.flash_attention(q1, k1, v1, True); O1.backward(dO.cuda())
    for i, (name, ref, one) in enumerate((("O", rO, O1), ("dQ", rdQ, q1.grad), ("dK", rdK, k1.grad), ("dV", rdV, v1.grad))):
        full = sh.zigzag_merge([o[i] for o in outs], dim=2).float()
        atol = 1e-2 * (math.sqrt(G) if name in ("dK", "dV") else 1.0)
        assert torch.allclose(full, ref.float(), atol=atol, rtol=1e-2), (name, (full - ref.float()).abs().max().item())
        assert torch.allclose(full, one.detach().cpu().float(), atol=2 * atol, rtol=2e-2), name


def test_virtual_ring_c5_head_vs_one_shot_and_truth():
    """One head of BASELINE config C5 (N = 131072, D = 128, bf16, causal) as a ring of 8 simulated ranks on one GPU against the
    one-shot kernel on the whole sequence, and both against fp32 truth on every element (contract tolerance)."""
    import flashattn_b200 as fa
    import flashattn_b200.sharding as sh
    from test_gpu_parity import _gpu_truth, _norm_err, _report
    world, N, D = 8, 131072, 128
    g = torch.Generator(device="cuda").manual_seed(77)
    Q, K, V, dO = (torch.randn(1, 1, N, D, device="cuda", generator=g).bfloat16() for _ in range(4))

    def rank_fn(rank, comm):
        torch.cuda.set_device(0)
        q, k, v, do = (sh.zigzag_split(t, rank, world) for t in (Q, K, V, dO))
        O, LSE = sh.ring_attention_forward(q, k, v, None, None, comm, 1)
        return (O,) + tuple(sh.ring_attention_backward(q, k, v, O, do, LSE, None, None, comm, 1))

    outs = sh.run_virtual_ring(world, rank_fn)
    q1, k1, v1 = (t.clone().requires_grad_(True) for t in (Q, K, V))
    O1 = fa.flash_attention(q1, k1, v1, True); O1.backward(dO)
    tO, _, tdQ, tdK, tdV = _gpu_truth(Q, K, V, dO, True)
    errs = {}
    for i, (name, t, one) in enumerate((("O", tO, O1), ("dQ", tdQ, q1.grad), ("dK", tdK, k1.grad), ("dV", tdV, v1.grad))):
        ring = sh.zigzag_merge([o[i] for o in outs], dim=2)
        errs[name] = dict(ring=_norm_err(ring, t), one_shot=_norm_err(one.detach(), t))
    _report("C5_head_ring8_vs_one_shot", max_norm_err=errs)
    assert all(max(e.values()) <= 1.0 for e in errs.values()), errs
