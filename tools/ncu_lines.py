"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep.
usage: python tools/ncu_lines.py rep.ncu-rep <kernel-name-regex> [min_pct]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kern}",
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
recs, cur, ix = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        ix = {k: j for j, k in enumerate(r)}
    elif len(r) > 10 and r[0].isdigit() and ix:
        try:
            recs.append((cur, int(r[0]), r[1], int(r[ix["Instructions Executed"]] or 0), int(r[ix["# Samples"]] or 0)))
        except ValueError:
            pass
ti, ts = sum(x[3] for x in recs) or 1, sum(x[4] for x in recs) or 1
print("instructions", ti, "samples", ts)
for f, l, src, i, s in recs:
    if 100.0 * i / ti >= min_pct or 100.0 * s / ts >= min_pct:
        print(f"{f}:{l:4d} inst={100.0 * i / ti:5.2f}% smp={100.0 * s / ts:5.2f}%  {src.strip()[:120]}")
