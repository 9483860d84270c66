"""One shape through fwd + fused backward a few times (for ncu captures).  python scripts/prof_fused.py B H S causal [D]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from flashattn_b200 import interface as I
B, H, S, causal = (int(x) for x in sys.argv[1:5]); causal = bool(causal); D = int(sys.argv[5]) if len(sys.argv) > 5 else 64
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V, dO = (torch.randn(B, H, S, D, device="cuda", generator=g).bfloat16() for _ in range(4))
dQ = torch.empty_like(Q); dK = torch.empty_like(K); dV = torch.empty_like(V)
delta = torch.empty(B, H, S, dtype=torch.float32, device="cuda"); acc = torch.empty(B, H, S, D, device="cuda")
for _ in range(3):
    O, LSE = I.flash_attention_forward(Q, K, V, causal)
    I.flash_attention_backward_fused(Q, K, V, O, dO, LSE, dQ, dK, dV, delta, causal, dq_acc=acc)
torch.cuda.synchronize()
print("ok")
