"""Per-instruction shared-memory wavefronts vs ideal from an `ncu --page source --csv` dump:
which LDS / STS cause the bank conflicts.   usage: ncu_conflicts.py file.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
recs = []; tot = ideal = 0
for r in rows[2:]:
    if len(r) >= 2 and r[0] == "Address" and r[1] == "Source": break
    if len(r) < len(hdr): continue
    w = float(r[ix["L1 Wavefronts Shared"]] or 0); i = float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    if w == 0: continue
    tot += w; ideal += i
    recs.append((w - i, w, i, int(float(r[ix["Instructions Executed"]] or 0)), r[ix["Address"]][-5:], r[ix["Source"]].strip()[:70]))
print(f"shared wavefronts {tot:.0f}, ideal {ideal:.0f}, excess {100 * (tot - ideal) / tot:.1f} %")
for e, w, i, n, a, s in sorted(recs, reverse=True)[:top]:
    print(f"  excess {e:10.0f} ({100 * e / tot:4.1f}%)  wavefronts/inst {w / max(n, 1):5.2f} ideal {i / max(n, 1):5.2f}  x{n:8d} {a} {s}")
