"""Executed-instruction histogram (by opcode, and the hottest basic stretches) of one kernel in an .ncu-rep.
usage: python tools/ncu_sass.py rep.ncu-rep <kernel-name-regex> [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
ix = {k: j for j, k in enumerate(rows[hdr])}
ins = []
for r in rows[hdr + 1:]:
    if r and r[0] in ("Address", "Kernel Name"):
        if ins: break
        continue
    if len(r) > ix["Instructions Executed"]:
        ins.append((r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
tot = sum(x[1] for x in ins)
print("instructions executed", tot, "static", len(ins))
ops = defaultdict(int)
for s, n, _ in ins:
    t = s.split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += n
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:top]:
    print(f"{k:10s} {100.0 * v / tot:6.2f}%")
# stretches of consecutive instructions with the same execution count
i = 0
seg = []
while i < len(ins):
    j = i
    while j + 1 < len(ins) and ins[j + 1][1] == ins[i][1]:
        j += 1
    seg.append((i, j, ins[i][1] * (j - i + 1)))
    i = j + 1
print("hottest stretches (first..last static index, executions each, share):")
for a, b, w in sorted(seg, key=lambda x: -x[2])[:top]:
    print(f"  {a:5d}..{b:5d} x{ins[a][1]:9d} {100.0 * w / tot:6.2f}%   {ins[a][0][:60]}")
