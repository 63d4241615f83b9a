"""Extracts the SASS of one kernel from an object file into profiles/ (gzip) and prints the mnemonic counts
that show what the kernel is made of.   usage: dump_sass.py <obj> <mangled-name-substring> <out.txt.gz>"""
import gzip, re, subprocess, sys
from collections import Counter
obj, pat, out = sys.argv[1:4]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = txt.split("\t\tFunction : ")
sel = [b for b in blocks[1:] if pat in b.split("\n", 1)[0]]
assert sel, "kernel not found"
body = "Function : " + sel[0]
gzip.open(out, "wt").write(body)
ops = Counter(m.group(1).split(".")[0] for m in re.finditer(r"\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]+)", body))
print(sel[0].split("\n", 1)[0][:90], "->", out)
print("  ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(24)))
print("   TMA/async:", {k: v for k, v in ops.items() if k in ("UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "LDGSTS", "FADD2", "FMUL2", "FFMA2")})
