"""Drop-in module name: `from My_FlashAttention_optimized import flash_attention` keeps working when this
package directory precedes the reference's code/ on PYTHONPATH (route A)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from flashattn_b200.interface import (FlashAttentionFunction, attention, flash_attention,  # noqa: E402,F401
                                      flash_attention_backward, flash_attention_forward)
