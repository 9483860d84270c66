"""Build the sm_100a CUDA library (and the bring-up binary) in-tree with nvcc.

    python flashattention-from-scratch-with-triton_b200/build.py [--force] [--bringup] [-v]

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    flashattention-from-scratch-with-triton_b200/libfa_sm100.so
    build/fa_bringup
nvcc cross-compiles for sm_100a without a GPU.  -lineinfo keeps the ncu source page usable.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libfa_sm100.so")
BRINGUP = os.path.join(ROOT, "build", "fa_bringup")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "--use_fast_math"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libfa_sm100.so cannot be built (there is no non-CUDA fallback)")
    return nvcc


def _stale(target: str, srcs: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _sources() -> list[str]:
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    out.append(os.path.join(ROOT, "include", "fa_sm100.h"))
    return out


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale(LIB, _sources()):
        return LIB
    cmd = [_nvcc(), *ARCH, *COMMON, "-shared", "-Xcompiler", "-fPIC", "-o", LIB, os.path.join(CSRC, "fa_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libfa_sm100.so")
    return LIB


def build_variant(name: str, defines: list[str]) -> str:
    """A/B builds for tuning: build/variants/libfa_sm100_<name>.so with extra -D flags; select at run time
    with FA_SM100_LIB=<path> (see _cabi.py)."""
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libfa_sm100_{name}.so")
    cmd = [_nvcc(), *ARCH, *COMMON, "-shared", "-Xcompiler", "-fPIC", *[f"-D{d}" for d in defines], "-o", out,
           os.path.join(CSRC, "fa_api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed building variant {name}")
    return out


def build_bringup(force: bool = False) -> str:
    os.makedirs(os.path.dirname(BRINGUP), exist_ok=True)
    srcs = [os.path.join(CSRC, "fa_bringup.cu"), os.path.join(CSRC, "fa_ptx.cuh")]
    if not force and not _stale(BRINGUP, srcs):
        return BRINGUP
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-o", BRINGUP, srcs[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building fa_bringup")
    return BRINGUP


TORCH_EXT = os.path.join(PKG, "_fa_torch.so")


def build_torch_ext(force: bool = False, verbose: bool = False) -> str:
    """_fa_torch.so: the C++ autograd host path (csrc/fa_torch.cpp) — host code only, compiled with g++ against the torch
    headers and linked to libfa_sm100.so next to it (rpath $ORIGIN).  In-tree, so it travels to the GPU box."""
    import sysconfig
    import torch
    from torch.utils import cpp_extension as ce
    build_lib()
    src = os.path.join(CSRC, "fa_torch.cpp")
    if not force and not _stale(TORCH_EXT, [src, os.path.join(ROOT, "include", "fa_sm100.h")]):
        return TORCH_EXT
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DTORCH_EXTENSION_NAME=_fa_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *[f"-I{d}" for d in ce.include_paths()], f"-I{sysconfig.get_paths()['include']}", f"-I{cuda_inc}",
           src, "-o", TORCH_EXT, f"-L{PKG}", "-l:libfa_sm100.so", f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
           "-lc10_cuda", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}", "-Wno-deprecated-declarations"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("g++ failed building _fa_torch.so")
    return TORCH_EXT


MICROBENCH = os.path.join(ROOT, "build", "fa_microbench")


def build_microbench(force: bool = False) -> str:
    """build/fa_microbench: tcgen05.ld / MUFU throughput probes (profiles/r01_microbench_tmem_mufu.txt)."""
    os.makedirs(os.path.dirname(MICROBENCH), exist_ok=True)
    srcs = [os.path.join(CSRC, "fa_microbench.cu"), os.path.join(CSRC, "fa_ptx.cuh")]
    if not force and not _stale(MICROBENCH, srcs):
        return MICROBENCH
    cmd = [_nvcc(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-o", MICROBENCH, srcs[0]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building fa_microbench")
    return MICROBENCH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--bringup", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--variant", action="append", default=[], help="name:DEF1=V,DEF2=V  (A/B build)")
    a = ap.parse_args()
    for v in a.variant:
        name, _, defs = v.partition(":")
        print(build_variant(name, [d for d in defs.split(",") if d]))
    print(build_lib(a.force, a.verbose))
    print(build_torch_ext(a.force, a.verbose))
    if a.bringup:
        print(build_bringup(a.force))
        print(build_microbench(a.force))
