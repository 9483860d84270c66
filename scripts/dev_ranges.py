import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import flashattn_b200 as fa
from oracle import attention_oracle as orc
g = torch.Generator().manual_seed(33)
B, H, S, D = 2, 3, 640, 64
Q, K, V, dO = (torch.randn(B, H, S, D, generator=g).bfloat16() for _ in range(4))
for name, r, causal in (("pad", fa.Ranges.from_key_padding([200, S], S, S, device="cuda"), False),
                        ("win", fa.Ranges.sliding_window(B, S, 200, device="cuda"), True)):
    O, LSE = fa.flash_attention_forward(Q.cuda(), K.cuda(), V.cuda(), causal, ranges=r)
    rO, rLSE = orc.closed_form(Q, K, V, None, causal, row_ranges=(r.row_lo.cpu(), r.row_hi.cpu()))
    d = (LSE.cpu() - rLSE.float()).abs()
    bad = torch.nonzero(~(d < 1e-3))
    print(name, "bad rows", bad.shape[0], bad[:8].tolist(), [(LSE.cpu()[tuple(i)].item(), rLSE[tuple(i)].item()) for i in bad[:8]])
    print(name, "O err", (O.cpu().float() - rO.float()).abs().max().item())
