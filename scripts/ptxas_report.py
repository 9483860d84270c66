"""Per-kernel registers / spills / stack from `nvcc -Xptxas -v` (stderr of build.py -v): python build.py --force -v 2>&1 | python scripts/ptxas_report.py"""
import re, subprocess, sys
txt = sys.stdin.read()
name = None
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void fa::", "")
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and name:
        spill = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", ln)
    if m and name:
        print(f"{name:60s} regs={m.group(1):>3s} stack={spill[0]:>3s} spill_st={spill[1]:>3s} spill_ld={spill[2]:>3s}")
        name = None
