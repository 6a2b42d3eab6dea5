"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, share, average."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    name = r[ki].split('(')[0][:70]
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    if r[ui] == 'us': v *= 1e3
    elif r[ui] == 'ms': v *= 1e6
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{k:72s} n={v[0]:5d} sum={v[1] / 1e6:9.3f} ms share={v[1] / tot:6.1%} avg={v[1] / v[0] / 1e3:8.1f} us")
