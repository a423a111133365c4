"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_shares.py list.csv"""
import csv, re, sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
    name = re.sub(r"\(.*$", "", r["Kernel Name"])
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
for name in sorted(tot, key=lambda k: -tot[k]):
    print(f"{name[:70]:70s} n={cnt[name]:4d} total={tot[name]:10.1f} us share={100 * tot[name] / s:5.1f}%")
