"""Per-rank compute of config C5 (B=1 H=32 N=131072 D=128 causal, P=8 zigzag) on ONE GPU, no communication:
   gather variant  = local queries [1,32,16384,128] x all keys [1,32,131072,128] with the zigzag Ranges, head groups of 8
   ring variant    = the 8 hops' kernels of rank r (local causal, 'all q x first half kv', 'second half q x all kv') + merges
Prints ms per step (fwd+bwd) and the per-GPU TFLOPS a perfectly overlapped 8-GPU run would reach."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
import flashattn_b200.sharding as sh
P, B, H, N, D = 8, 1, 32, 131072, 128
rank = int(sys.argv[1]) if len(sys.argv) > 1 else 3
groups = int(sys.argv[2]) if len(sys.argv) > 2 else 4
c = N // (2 * P); S2 = 2 * c
g = torch.Generator(device="cuda").manual_seed(rank)
q, do = (torch.randn(B, H, S2, D, device="cuda", generator=g).bfloat16() for _ in range(2))
ev = lambda: torch.cuda.Event(enable_timing=True)
flops_rank = 3.5 * 4 * B * H * N * N * D / 2 / P
out = {"rank": rank, "groups": groups}
# ---------------- gather variant: one range-masked launch per head group
ranges = sh.zigzag_ranges(rank, P, c, B, "cuda")
hg = H // groups
Kg, Vg = (torch.randn(B, hg, N, D, device="cuda", generator=g).bfloat16() for _ in range(2))     # one group's K/V, reused for every group
ops = sh.GatherOps()
def gather_step():
    for gi in range(groups):
        qs = slice(gi * hg, (gi + 1) * hg)
        O, L = ops.fwd(q[:, qs], Kg, Vg, ranges)
        ops.bwd(q[:, qs], Kg, Vg, O, do[:, qs], L, ranges)
for _ in range(2): gather_step()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    s, e = ev(), ev(); s.record(); gather_step(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
out["gather_ms"] = round(min(ts), 3); out["gather_tflops_per_gpu"] = round(flops_rank / min(ts) / 1e9, 1)
# forward / backward split
s, m, e = ev(), ev(), ev()
s.record(); Os = [ops.fwd(q[:, gi * hg:(gi + 1) * hg], Kg, Vg, ranges) for gi in range(groups)]; m.record()
for gi in range(groups): ops.bwd(q[:, gi * hg:(gi + 1) * hg], Kg, Vg, Os[gi][0], do[:, gi * hg:(gi + 1) * hg], Os[gi][1], ranges)
e.record(); torch.cuda.synchronize()
out["gather_fwd_ms"] = round(s.elapsed_time(m), 3); out["gather_bwd_ms"] = round(m.elapsed_time(e), 3)
del Kg, Vg, Os
# ---------------- ring variant: rank r's P hops (compute + merges only)
k, v = (torch.randn(B, H, S2, D, device="cuda", generator=g).bfloat16() for _ in range(2))
rops = sh.CudaOps()
def ring_step():
    O_acc = torch.empty(B, H, S2, D, dtype=torch.float32, device="cuda"); L_acc = torch.empty(B, H, S2, dtype=torch.float32, device="cuda")
    for s_ in range(P):
        o = (rank - s_) % P
        if o == rank:
            Op, Lp = rops.fwd(q, k, v, True); O_acc.copy_(Op); L_acc.copy_(Lp)
        elif o < rank:
            Op, Lp = rops.fwd(q, k[:, :, :c], v[:, :, :c], False); rops.merge_(O_acc, L_acc, Op, Lp, 0)
        else:
            Op, Lp = rops.fwd(q[:, :, c:], k, v, False); rops.merge_(O_acc, L_acc, Op, Lp, c)
    O = O_acc.to(q.dtype)
    delta = rops.delta(O, do)
    L_hi, d_hi = L_acc[:, :, c:].contiguous(), delta[:, :, c:].contiguous()
    dq_acc = torch.zeros(B, H, S2, D, dtype=torch.float32, device="cuda")
    acc = [torch.zeros(B, H, S2, D, dtype=torch.float32, device="cuda") for _ in range(2)]
    for s_ in range(P):
        o = (rank - s_) % P
        if o == rank:
            dq, dk, dv = rops.bwd(q, k, v, O, do, L_acc, delta, True); dq_acc.add_(dq); acc[0].add_(dk); acc[1].add_(dv)
        elif o < rank:
            dq, dk, dv = rops.bwd(q, k[:, :, :c], v[:, :, :c], O, do, L_acc, delta, False); dq_acc.add_(dq); acc[0][:, :, :c].add_(dk); acc[1][:, :, :c].add_(dv)
        else:
            dq, dk, dv = rops.bwd(q[:, :, c:], k, v, O[:, :, c:], do[:, :, c:], L_hi, d_hi, False); dq_acc[:, :, c:].add_(dq); acc[0].add_(dk); acc[1].add_(dv)
for _ in range(2): ring_step()
torch.cuda.synchronize()
ts = []
for _ in range(3):
    s, e = ev(), ev(); s.record(); ring_step(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
out["ring_compute_ms"] = round(min(ts), 3); out["ring_tflops_per_gpu_if_comm_hidden"] = round(flops_rank / min(ts) / 1e9, 1)
print(json.dumps(out), flush=True)
