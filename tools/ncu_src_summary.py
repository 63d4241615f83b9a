"""Summarise an `ncu --page source --csv` dump: executed warp instructions and stall samples
per opcode and the top stalled instructions.  usage: ncu_src_summary.py file.csv [topN]"""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ops, samples, stalls = Counter(), Counter(), Counter()
tot_inst = tot_samp = 0
recs = []
table = []
for r in rows[2:]:
    if len(r) >= 2 and r[0] == "Address" and r[1] == "Source":
        break                                   # a second view (source lines) follows the SASS view
    table.append(r)
for r in table:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0]
    ne = int(float(r[ix["Instructions Executed"]] or 0))
    ns = int(float(r[ix["# Samples"]] or 0))
    ops[op] += ne
    samples[op] += ns
    tot_inst += ne
    tot_samp += ns
    for s in stall_cols:
        v = r[ix[s]]
        if v:
            stalls[s] += int(float(v))
    recs.append((ns, ne, r[ix["Address"]], src[:90], {s: int(float(r[ix[s]] or 0)) for s in stall_cols}))
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
print("stall reasons:", ", ".join(f"{k[6:]}={v * 100.0 / max(1, sum(stalls.values())):.1f}%" for k, v in stalls.most_common(8)))
print("by opcode (executed, share, samples share):")
for op, n in ops.most_common(22):
    print(f"  {op:10s} {n:12d} {n * 100.0 / tot_inst:5.1f}%   samples {samples[op] * 100.0 / max(1, tot_samp):5.1f}%")
print("top stalled instructions:")
for ns, ne, addr, src, st in sorted(recs, reverse=True)[:top]:
    main = max(st, key=st.get)
    print(f"  {ns:6d} ({ns * 100.0 / tot_samp:4.1f}%) x{ne:9d} {addr[-5:]} {src:70s} {main[6:]}")
