"""Per-source-line warp-stall samples of one kernel of an .ncu-rep (dev tool):
   python tools/ncu_stalls.py rep.ncu-rep KERNEL_ID [top]"""
import collections
import csv
import io
import subprocess
import re
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        cur = {"file": row[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif row[0] == "Function Name":
        continue
    elif row[0] == "Line No":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
stall_cols = None
tot = collections.Counter()
per_line = []
inst_tot = 0
ops = collections.Counter()
for b in blocks:
    h = b["hdr"]
    si = h.index("# Samples")
    ii = h.index("Instructions Executed")
    names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    idx = {n: h.index(n) for n in names}
    for r in b["rows"]:
        if r[0] == "":          # SASS row under the source line: opcode histogram
            try:
                n_i = int(r[ii] or 0)
            except ValueError:
                continue
            m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[3])
            if m:
                ops[m.group(2).split(".")[0]] += n_i
            continue
        try:
            s = int(r[si] or 0)
            n_inst = int(r[ii] or 0)
        except ValueError:
            continue
        inst_tot += n_inst
        st = {n: int(r[i] or 0) for n, i in idx.items() if r[i] not in ("", "0")}
        for n, v in st.items():
            tot[n] += v
        if s or n_inst:
            per_line.append((s, n_inst, b["file"].split("/")[-1], r[0], r[1].strip()[:90], st))
all_s = sum(x[0] for x in per_line)
print(f"samples {all_s}  warp-instructions {inst_tot}")
print("opcodes:", ", ".join(f"{k}={v}" for k, v in ops.most_common(28)))
print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in tot.most_common(10)))
for s, n_inst, f, ln, src, st in sorted(per_line, key=lambda x: -x[0])[:top]:
    main = ",".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:7d} {100.0 * s / max(all_s, 1):5.1f}% inst={n_inst:9d} {f}:{ln:>4} {src}   [{main}]")
print("--- by instruction count")
for s, n_inst, f, ln, src, st in sorted(per_line, key=lambda x: -x[1])[:top]:
    print(f"inst={n_inst:9d} {100.0 * n_inst / max(inst_tot, 1):5.1f}% samples={s:6d} {f}:{ln:>4} {src}")
