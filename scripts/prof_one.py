"""Run each kernel a few times on one shape (for ncu captures).  python scripts/prof_one.py B H S D causal"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
B, H, S, D, causal = (int(x) for x in sys.argv[1:6]); causal = bool(causal)
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(4))
dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V)
delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
for _ in range(3):
    O, LSE = fa.flash_attention_forward(Q, K, V, causal)
    fa.flash_attention_backward_parts(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, 7)
torch.cuda.synchronize()
print("ok")
