#!/usr/bin/env python3
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (read on the CPU box).
   python tools/launch_summary.py launches.csv ["header comment"]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
t, n = collections.Counter(), collections.Counter()
for r in rows[hdr + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[kn])
    t[name] += float(r[mv].replace(",", "")) / 1e3
    n[name] += 1
own = {k: v for k, v in t.items() if re.search(r"\bk_[a-z]", k)}
tot = sum(own.values())
for c in sys.argv[2:]:
    print("# " + c)
for k, v in sorted(own.items(), key=lambda kv: -kv[1]):
    print("%-60s n=%5d %12.1f us %6.1f%%" % (k[:60], n[k], v, 100 * v / tot))
oth = sum(v for k, v in t.items() if k not in own)
print("# kernels of this library: %.1f us in %d launches; torch / NCCL kernels in the same run: %.1f us in %d launches" %
      (tot, sum(n[k] for k in own), oth, sum(n[k] for k in t if k not in own)))
