"""B200-native FlashAttention behind the reference's PyTorch operator API.

Importable as ``flashattn_b200`` (shim at the repo root; this directory's name is not a valid
Python identifier).  Hot path: hand-written sm_100a CUDA in ``csrc/`` behind the C ABI in
``include/fa_sm100.h``; this package is the host-side mirror of the reference operator.
"""
from .interface import (FlashAttentionFunction, attention, flash_attention, flash_attention_bshd, tma_compatible, flash_attention_backward, flash_attention_backward_parts,
                        flash_attention_delta, flash_attention_forward, merge_partial_, flash_attention_backward_fused,
                        set_deterministic, is_deterministic, set_fused128, set_shared_sms, Ranges, flash_attention_varlen)
from .host_pipeline import HostAttentionPipeline, flash_attention_host
from .verify import verify_results
from .flops import attention_flops, tflops

__all__ = ["flash_attention", "attention", "flash_attention_bshd", "tma_compatible", "FlashAttentionFunction", "flash_attention_forward",
           "flash_attention_backward", "flash_attention_backward_parts", "flash_attention_delta", "merge_partial_", "verify_results",
           "attention_flops", "tflops", "HostAttentionPipeline", "flash_attention_host", "flash_attention_backward_fused",
           "set_deterministic", "is_deterministic", "set_fused128", "set_shared_sms", "Ranges", "flash_attention_varlen"]
