"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` output: stall reasons and hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# first kernel instance only
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]; data = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    if len(r) == len(hdr): data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
S = lambda r: int(r[ix["# Samples"]])
tot = sum(S(r) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {c: sum(int(r[ix[c]]) for r in data) for c in stall_cols}
for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]: print(f"{c:28s} {v:7d} {100*v/tot:5.1f}%")
for r in sorted(data, key=lambda r: -S(r))[:ntop]:
    reasons = sorted(((int(r[ix[c]]), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{S(r):6d} {100*S(r)/tot:4.1f}% exec={r[ix['Instructions Executed']]:>8s} {r[ix['Address']][-5:]} {r[ix['Source']].strip()[:60]:60s} {reasons}")
