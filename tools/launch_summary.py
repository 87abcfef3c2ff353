"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / total / average."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for x in r:
    name = x[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:48]
    v = float(x[vi].replace(',', ''))
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
    if len(sys.argv) > 2 and sys.argv[2] in x[ki]:
        print(f"  {name:40s} grid {x[gi]:16s} {v / 1e3:8.2f} us")
print(f"{sum(n for n, _ in agg.values())} launches, {tot / 1e3:.1f} us in kernels")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:50s} n={n:4d} total={t / 1e3:9.1f} us  avg={t / n / 1e3:7.2f} us  {100 * t / tot:5.1f} %")
