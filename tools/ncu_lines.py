"""Per-source-line view of an .ncu-rep (needs -lineinfo + --import-source on): executed warp instructions and
stall samples per CUDA source line.   usage: ncu_lines.py file.ncu-rep kernel_regex [topN]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, recs, hdr = None, [], None
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i - len(r) for i, h in enumerate(r) if h != "Source"}       # from the end: source text may hold commas / quotes
    elif hdr and len(r) > 8 and r[0].isdigit():
        try:
            ne = int(float(r[hdr["Instructions Executed"]] or 0))
            ns = int(float(r[hdr["# Samples"]] or 0))
        except ValueError:
            continue
        if ne or ns:
            recs.append((ne, ns, cur, int(r[0]), ",".join(r[1:len(r) + hdr["Address"]]).strip()[:110]))
ti, ts = sum(r[0] for r in recs), sum(r[1] for r in recs)
print(f"total warp instructions {ti}, samples {ts}")
for ne, ns, f, ln, src in sorted(recs, reverse=True)[:top]:
    print(f"{ne * 100.0 / ti:5.1f}% inst {ns * 100.0 / max(1, ts):5.1f}% smp  {f}:{ln:<4d} {src}")
