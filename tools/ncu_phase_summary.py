"""Splits the SASS of a kernel (ncu --page source --csv dump) at its BAR.SYNC instructions and
prints, per region, executed warp instructions, stall samples and the dominant stall reasons.
usage: ncu_phase_summary.py file.csv"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
regions = [dict(n=0, s=0, st=Counter(), first=None, ops=Counter())]
for r in rows[2:]:
    if len(r) >= 2 and r[0] == "Address" and r[1] == "Source": break
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    reg = regions[-1]
    ne = int(float(r[ix["Instructions Executed"]] or 0)); ns = int(float(r[ix["# Samples"]] or 0))
    if reg["first"] is None: reg["first"] = r[ix["Address"]][-5:]
    reg["n"] += ne; reg["s"] += ns
    op = src.split()[1] if src.startswith("@") else (src.split()[0] if src else "?")
    reg["ops"][op.split(".")[0]] += ne
    for s in stall_cols:
        if r[ix[s]]: reg["st"][s[6:]] += int(float(r[ix[s]]))
    if "BAR.SYNC" in src:
        regions.append(dict(n=0, s=0, st=Counter(), first=None, ops=Counter()))
tn = sum(r["n"] for r in regions); ts = sum(r["s"] for r in regions)
print(f"total warp instructions {tn}, samples {ts}")
for i, r in enumerate(regions):
    if r["n"] == 0: continue
    tot = max(1, sum(r["st"].values()))
    print(f"region {i} @{r['first']}: inst {r['n']*100/tn:5.1f}%  samples {r['s']*100/ts:5.1f}%  cyc/inst/warp {r['s']/max(1,r['n'])*tn/ts:5.2f}x  | "
          + ", ".join(f"{k}={v*100/tot:.0f}%" for k, v in r["st"].most_common(5))
          + " | " + ", ".join(f"{k}:{v*100/r['n']:.0f}%" for k, v in r["ops"].most_common(6)))
