"""The C-ABI library loads, exports every symbol include/fa_sm100.h declares, and rejects bad
arguments with the documented codes — all without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fa_sm100.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    d = _declared()
    for name in ("fa_sm100_fwd", "fa_sm100_bwd", "fa_sm100_bwd_parts", "fa_sm100_fwd_strided", "fa_sm100_bwd_strided", "fa_sm100_delta", "fa_sm100_merge", "fa_sm100_supported",
                 "fa_last_error", "fa_sm100_version", "fa_sm100_launch_count", "fa_sm100_last_hang"):
        assert name in d


def test_library_exports_every_declared_symbol(lib):
    import flashattn_b200._cabi as cabi
    assert sorted(cabi.SYMBOLS) == _declared()
    for name in _declared():
        assert getattr(lib, name) is not None
    assert lib.fa_sm100_version() >= 100


def test_supported_query(lib):
    assert lib.fa_sm100_supported(64, 0, 128, 128) == 1
    assert lib.fa_sm100_supported(128, 1, 131072, 131072) == 1
    assert lib.fa_sm100_supported(96, 1, 128, 128) == 0
    assert lib.fa_sm100_supported(64, 2, 128, 128) == 0
    assert lib.fa_sm100_supported(64, 1, 0, 128) == 0


def test_argument_validation_without_gpu(lib):
    p = 0x10000     # fake, aligned, never dereferenced: validation happens before any CUDA call
    fwd = lambda q=p, D=64, dt=1, Sq=128, Sk=128, B=1: lib.fa_sm100_fwd(q, p, p, p, p, B, 2, Sq, Sk, D, dt, 0, 0.0, None)
    assert fwd(q=None) == -1 and b"null" in lib.fa_last_error()
    assert fwd(dt=3) == -2
    assert fwd(D=96) == -3 and b"head dim" in lib.fa_last_error()
    assert fwd(Sq=0) == -4
    assert fwd(B=-1) == -4
    assert fwd(q=p + 2) == -5
    bwd = lambda dq=p, D=128, dt=0: lib.fa_sm100_bwd(p, p, p, p, p, p, dq, p, p, p, 1, 1, 128, 128, D, dt, 1, 0.0, None)
    assert bwd(dq=None) == -1
    assert bwd(D=32) == -3
    assert bwd(dt=-1) == -2
    assert bwd(dq=p + 8) == -5
    import ctypes
    bad = (ctypes.c_longlong * 12)(*([128 * 64 * 2, 128 * 64, 64] * 3 + [128 * 64 * 2, 128 * 64, 60]))   # o row stride 60: not 16-byte
    assert lib.fa_sm100_fwd_strided(p, p, p, p, p, 1, 2, 2, 128, 128, 64, 1, 0, 0.0, bad, None) == -8
    assert lib.fa_sm100_fwd_strided(p, p, p, p, p, 1, 6, 4, 128, 128, 64, 1, 0, 0.0, None, None) == -4   # Hk must divide H
    fused = lambda D=64, acc=p: lib.fa_sm100_bwd_fused(p, p, p, p, p, p, p, p, p, p, acc, 1, 2, 2, 128, 128, D, 1, 1, 0.0, None, None, 0)
    assert fused(D=96) == -3 and b"head dim" in lib.fa_last_error()        # fused kernels exist for head dims 64 and 128
    assert fused(acc=None) == -1
    assert lib.fa_sm100_bwd_fused_workspace(2, 4, 256, 64) == 2 * 4 * 256 * 64 * 4
    assert lib.fa_sm100_bwd_fused_workspace(2, 4, 256, 128) == 2 * 4 * 256 * 128 * 4 and lib.fa_sm100_bwd_fused_workspace(2, 4, 256, 96) == 0
    # range masks: lo / hi arrays come in pairs (quadruples for the backward)
    assert lib.fa_sm100_fwd_ranges(p, p, p, p, p, 1, 2, 2, 128, 128, 64, 1, 0, 0.0, None, p, None, None) == -1
    assert lib.fa_sm100_bwd_ranges(p, p, p, p, p, p, p, p, p, p, 1, 2, 2, 128, 128, 64, 1, 0, 0.0, None, p, p, p, None, None, 7) == -1
    # options struct: dropout_p must be in [0, 1)
    import flashattn_b200._cabi as cabi
    opt = cabi.Options(); opt.dropout_p = 1.0
    assert lib.fa_sm100_fwd_opt(p, p, p, p, p, 1, 2, 2, 128, 128, 64, 1, 0, 0.0, None, ctypes.byref(opt), None) == -4
    assert b"dropout" in lib.fa_last_error()
    assert lib.fa_sm100_delta(p, None, p, 1, 1, 128, 64, 1, None) == -1
    assert lib.fa_sm100_merge(p, p, p, None, 1, 1, 128, 64, 1, 128, 0, None) == -1
    assert lib.fa_sm100_merge(p, p, p, p, 1, 1, 128, 64, 1, 128, 64, None) == -4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import flashattn_b200._cabi as cabi
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU or Triton fallback"):
        cabi.load()


def test_operator_asserts_like_the_reference():
    """reference code/My_FlashAttention_optimized.py:133-136 — AssertionError, not a fallback."""
    import torch
    import flashattn_b200 as fa
    q = torch.randn(1, 1, 128, 64, dtype=torch.float16)
    with pytest.raises(AssertionError):
        fa.flash_attention(q, q, q)                              # CPU tensors
    with pytest.raises(AssertionError):
        fa.FlashAttentionFunction.apply(q.float(), q.float(), q.float(), False)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, open or link it."""
    pkg = os.path.join(ROOT, "flashattention-from-scratch-with-triton_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]attention_oracle|oracle/_ref", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f


def test_header_is_plain_c_and_links_against_the_library(tmp_path, lib):
    """include/fa_sm100.h is the drop-in boundary for ANY host language: it must compile as C99 (no C++, no torch types) and a C
    program linked against libfa_sm100.so must resolve every entry point it declares (called here only for argument validation
    and the capability query: no GPU needed)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "flashattention-from-scratch-with-triton_b200")
    src = tmp_path / "use_header.c"
    src.write_text('#include <stdio.h>\n#include "fa_sm100.h"\n'
                   "int main(void) {\n"
                   "  int rc = fa_sm100_fwd(0, 0, 0, 0, 0, 1, 2, 128, 128, 64, 1, 0, 0.0f, 0);   /* null tensors: validation only */\n"
                   '  printf("%d %d %d %s\\n", fa_sm100_version(), fa_sm100_supported(128, 1, 77, 300), rc, fa_last_error());\n'
                   "  return 0;\n}\n")
    exe = tmp_path / "use_header"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L" + pkg, "-l:libfa_sm100.so", "-Wl,-rpath," + pkg], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split(maxsplit=3)
    assert int(out[0]) >= 100 and int(out[1]) == 1 and int(out[2]) == -1 and "null" in out[3]
