"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
cols, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = cols.index('Kernel Name'), cols.index('Metric Value'), cols.index('Metric Unit')
gi = cols.index('Grid Size') if 'Grid Size' in cols else None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    t = float(r[vi].replace(',', ''))
    t = t / 1e3 if r[ui] == 'ns' else (t * 1e3 if r[ui] == 'ms' else t)
    name = r[ki].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"# total {tot:.0f} us over {len(data)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.0f} us {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {k}")
