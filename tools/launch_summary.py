# SPDX-License-Identifier: Apache-2.0
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean and total
time per kernel name, share of the listed total.  python tools/launch_summary.py launches.csv [skip_regex]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = None
for i, r in enumerate(rows):
    if "Kernel Name" in r and "Metric Value" in r:
        hdr = i
        break
if hdr is None:
    sys.exit("no ncu csv header found")
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn])
    if skip and skip.search(name):
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    unit = r[mu]
    us = v / 1e3 if unit.startswith("ns") else (v if unit.startswith("us") else v * 1e3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name[:100]}` | {n} | {t / n:.1f} | {t:.0f} | {100 * t / tot:.1f}% |")
print(f"\ntotal {tot:.0f} us over {sum(a[0] for a in agg.values())} launches")
