"""Kernel shares from an ncu launch list (--metrics gpu__time_duration.sum --csv):
  python tests/probe/summarise_launches.py gpurun_out/launches.csv "header line" > profiles/rNN_launches_X.summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head = rows[0]
ki, vi, ui = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki]).replace("(anonymous namespace)::", "").strip()
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
for h in sys.argv[2:]:
    print("# " + h)
print("# total %.1f ms over %d launches" % (total, sum(cnt.values())))
for k in sorted(tot, key=tot.get, reverse=True)[:30]:
    print("%-64s %4d launches %9.3f ms %5.1f%%" % (k[:64], cnt[k], tot[k], 100 * tot[k] / total))
