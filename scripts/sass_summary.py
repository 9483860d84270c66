#!/usr/bin/env python
"""Opcode histogram per kernel of libfa_sm100.so (cuobjdump -sass): the mnemonics that prove the tcgen05 / TMEM / TMA path
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, MUFU.EX2, FFMA2 / FADD2 / FMUL2 packed fp32, FMNMX3, ACQBULK / USETMAXREG).

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "flashattention-from-scratch-with-triton_b200", "libfa_sm100.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "SYNCS", "MUFU.EX2", "MUFU.LG2", "MUFU.RCP",
        "FFMA2", "FADD2", "FMUL2", "FMNMX3", "FMNMX", "F2FP", "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "USETMAXREG", "ELECT", "BAR",
        "LDL", "STL", "ACQBULK"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    kernels[cur][k] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  ({len(kernels)} kernels, sm_100a)")
    print("# spills: LDL/STL counts; tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG")
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void fa::", "")
        items = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
        print(f"{short:60s} instr={c['_total']:6d}  {items}")
        tot.update(c)
    print("TOTAL".ljust(60), f"instr={tot['_total']:6d} ", " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
