"""profiles/launches_*.csv (ncu --metrics gpu__time_duration.sum --csv) -> markdown table of kernel
shares of one SFC forward.   python scripts/launch_shares.py profiles/launches_r01b.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("w2v::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"Total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches.\n")
print("| share | launches | total us | kernel |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {100 * t / tot:.2f}% | {n} | {t:.1f} | `{k}` |")
