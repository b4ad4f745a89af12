"""Opcode histogram (executed warp-instructions) + stall samples by opcode of the first kernel in an .ncu-rep (dev tool):
   python tools/ncu_sass_hist.py rep.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = next(r for r in rows if r and r[0] == "Address")
si, ii, smp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, stall = collections.Counter(), collections.Counter()
total = 0
lines = []
for r in rows:
    if len(r) <= ii or not r[0].startswith("0x"):
        continue
    src = r[si].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if "." in op and op.split(".")[0] in ("MUFU", "LDTM", "STTM", "F2FP", "SYNCS", "BAR") else "")
    n = int(r[ii] or 0)
    s = int(r[smp] or 0)
    ops[op] += n
    stall[op] += s
    total += n
    lines.append((s, n, src))
print(f"warp-instructions executed: {total}")
for op, n in ops.most_common(top):
    print(f"{n:10d} {100.0 * n / total:5.1f}%  samples {stall[op]:6d}  {op}")
print("--- top stall lines")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(f"{s:6d} samples {n:9d} exec  {src[:100]}")
