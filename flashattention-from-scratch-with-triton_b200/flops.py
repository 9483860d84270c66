"""The reference's FLOP model (code/Performance_Comparison.py:99-107), bit for bit."""


def attention_flops(B, H, S_q, S_k, D, is_causal, mode="fwd"):
    flops = 4 * B * H * S_q * S_k * D // (2 if is_causal else 1)      # :101
    if mode == "fwd":
        return flops                                                  # :103
    if mode == "bwd":
        return 2.5 * flops                                            # :105
    if mode == "fwd_bwd":
        return 3.5 * flops                                            # :107
    raise ValueError(mode)


def tflops(B, H, S_q, S_k, D, is_causal, mode, ms):
    return attention_flops(B, H, S_q, S_k, D, is_causal, mode) / (ms * 1e-3) / 1e12
