"""Drop-in module name: `from My_FlashAttention_optimized import flash_attention` keeps working when this
package directory precedes the reference's code/ on PYTHONPATH (route A), and
`python My_FlashAttention_optimized.py` runs the reference's own self-check (compare_with_sdpa, reference
code/My_FlashAttention_optimized.py:172-226) against this library instead of the Triton kernels."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from flashattn_b200.interface import (FlashAttentionFunction, attention, flash_attention,  # noqa: E402,F401
                                      flash_attention_backward, flash_attention_forward)
from flashattn_b200.verify import verify_results  # noqa: E402


def compare_with_sdpa(Q, K, V, is_causal, verbose=True):
    """The reference's self-check (:172-212): PyTorch SDPA (flash backend, fp16 autocast) as the yardstick, a shared random dO,
    then verify_results on O, dQ, dK, dV (rtol 1e-2, atol 1e-3, cosine > 0.999).  Returns the four result dicts."""
    import torch
    import torch.nn.functional as F
    from torch.amp import autocast
    from torch.nn.attention import SDPBackend, sdpa_kernel
    Q_ref, K_ref, V_ref = (t.detach().clone().requires_grad_(True) for t in (Q, K, V))
    with sdpa_kernel(SDPBackend.FLASH_ATTENTION):                            # :178
        with autocast(device_type="cuda", dtype=Q.dtype):                    # :179 (the reference hard-wires fp16)
            O_ref = F.scaled_dot_product_attention(Q_ref, K_ref, V_ref, attn_mask=None, dropout_p=0.0, is_causal=is_causal)
    dO = torch.randn_like(O_ref)                                             # :189
    O_ref.backward(dO)
    Q_, K_, V_ = (t.detach().clone().requires_grad_(True) for t in (Q, K, V))
    O = flash_attention(Q_, K_, V_, is_causal=is_causal)                     # :199
    O.backward(dO)
    out = []
    for name, ref, got in (("O", O_ref, O), ("dQ", Q_ref.grad, Q_.grad), ("dK", K_ref.grad, K_.grad), ("dV", V_ref.grad, V_.grad)):
        if verbose:
            print(f"{'=' * 30} {name} test {'=' * 30}")                       # :203-210
        out.append(verify_results(ref, got, name, verbose=verbose))
    return out


if __name__ == "__main__":                                                   # :214-226
    import torch
    DEVICE = torch.device(torch.cuda.current_device())
    B, H, S_q, S_k, D = 4, 8, 256, 256, 64
    Q = torch.randn((B, H, S_q, D), dtype=torch.float16, device=DEVICE)
    K = torch.randn((B, H, S_k, D), dtype=torch.float16, device=DEVICE)
    V = torch.randn((B, H, S_k, D), dtype=torch.float16, device=DEVICE)
    res = compare_with_sdpa(Q, K, V, is_causal=True)
    _sys.exit(0 if all(r["passed"] for r in res) else 1)
