"""One head group of the gather variant's forward at C5 / P = 8 size on one GPU (for ncu): local queries [1,8,16384,128] of rank 3
against all keys [1,8,131072,128] under the zigzag Ranges."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200.sharding as sh
P, rank, c = 8, 3, 8192
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(1, 8, 2 * c, 128, device="cuda", generator=g).bfloat16()
K, V = (torch.randn(1, 8, 2 * c * P, 128, device="cuda", generator=g).bfloat16() for _ in range(2))
r = sh.zigzag_ranges(rank, P, c, 1, "cuda")
ops = sh.GatherOps()
for _ in range(3):
    O, L = ops.fwd(q, K, V, r)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.fwd(q, K, V, r); e.record(); torch.cuda.synchronize()
fl = 4 * 8 * (2 * c) * (2 * c * P) * 128 / 2 * (2 * P + 1) / (2 * P) / P * P / 8   # ~ causal share of this rank
print("fwd ms", s.elapsed_time(e))
