"""Per-kernel summary of an ncu --set full report: python scripts/ncu_summary.py report.ncu-rep workload > summary.json
(duration, clock, DRAM bytes, tensor / XU / issue activity, shared-memory pipe wavefronts, registers, smem)."""
import csv, io, json, subprocess, sys
rep, workload = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]
ix = {h: i for i, h in enumerate(hdr)}
units = rows[1]
out, seen = [], set()
for r in rows[2:]:
    if len(r) != len(hdr): continue
    name = r[ix["Kernel Name"]].split("(")[0]
    if name in seen: continue
    seen.add(name)
    d = {"workload": workload, "kernel": name}
    for k in KEYS:
        if k in ix: d[k] = f"{r[ix[k]]} {units[ix[k]]}".strip()
    def num(k):
        v = float(r[ix[k]].replace(",", "")); u = units[ix[k]].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    d["dram_bytes_total"] = int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    out.append(d)
print(json.dumps(out, indent=1))
