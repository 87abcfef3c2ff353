#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a tracked profiles/ file:
per-kernel totals + (optionally) the per-launch rows of a window.
usage: launch_csv.py in.csv out.csv "header comment" [first_mlb_launch n_launches]   (window in units of mlb:: launches)"""
import collections, csv, sys
src, dst, note = sys.argv[1:4]
rows = list(csv.reader(open(src)))
i = next(k for k, r in enumerate(rows) if r and r[0] == 'ID')
hh = rows[i]; kn = hh.index('Kernel Name'); mv = hh.index('Metric Value')
seq = [(r[0], r[kn].split('(')[0].replace('void ', '').replace('(anonymous namespace)::', '').strip(), float(r[mv].replace(',', '')))
       for r in rows[i + 1:] if len(r) > mv]
agg = collections.OrderedDict()
for _, n, tm in seq: agg.setdefault(n, []).append(tm)
tot = sum(sum(v) for v in agg.values())
with open(dst, 'w') as fo:
    fo.write(f"# {note}\n# per-launch times under ncu are cold-cache and serialised: shares matter, not absolutes\n")
    fo.write(f"# {len(seq)} launches, {tot / 1e3:.1f} us in kernels\nkernel,launches,total_us,mean_us,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        fo.write(f"\"{k[:90]}\",{len(v)},{sum(v) / 1e3:.1f},{sum(v) / len(v) / 1e3:.2f},{sum(v) / tot:.3f}\n")
    if len(sys.argv) > 5:
        a, n = int(sys.argv[4]), int(sys.argv[5])
        mlb = [x for x in seq if 'mlb::' in x[1]][a:a + n]
        per = collections.OrderedDict()
        for _, k, tm in mlb: per.setdefault(k, []).append(tm)
        t2 = sum(sum(v) for v in per.values())
        fo.write(f"# window: mlb:: launches {a}..{a + n}: kernel,mean_us,share_of_step\n")
        for k, v in per.items(): fo.write(f"\"{k}\",{sum(v) / len(v) / 1e3:.1f},{sum(v) / t2:.3f}\n")
        fo.write("# per-launch rows of the window: id,kernel,duration_ns\n")
        for idn, k, tm in mlb: fo.write(f"{idn},\"{k}\",{tm:.0f}\n")
