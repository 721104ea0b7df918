#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: tools/launch_summary.py file.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= mv:
        continue
    name = r[kn].split("(")[0].replace("void ", "").replace("unnamed>::", "")
    agg[name][0] += 1
    agg[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot/1e3:.1f} us total (ncu-serialised, cold cache: compare SHARES)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={v[0]:5d} total={v[1]/1e3:10.1f} us  avg={v[1]/v[0]/1e3:9.2f} us  share={v[1]/tot*100:5.1f}%")
