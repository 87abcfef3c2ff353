#!/usr/bin/env python
"""Summarise the SASS page of an ncu report per kernel: runs of instructions with equal execution counts.
usage: sass_runs.py report.ncu-rep [min_total_instr] [kernel_substr] [a-b dump range]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 3e6
ksel = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for a, b in zip(starts, starts[1:]):
    name = rows[a][1]
    if name in seen or ksel not in name:
        continue
    seen.add(name)
    blk = rows[a:b]
    hdr = next(r for r in blk if "Instructions Executed" in r)
    iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = []
    for r in blk:
        try:
            data.append((r[1].strip(), int(r[iI]), int(r[iS])))
        except Exception:
            pass
    tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
    print("====", name[:70]); print("total warp instr", tot, "samples", ts, "static instrs", len(data))
    runs = []
    for i, (s, n, sm) in enumerate(data):
        if runs and runs[-1][1] == n:
            runs[-1][2] += 1; runs[-1][3] += sm; runs[-1][4] = i
        else:
            runs.append([i, n, 1, sm, i])
    for st, n, l, sm, en in runs:
        if n * l > thr:
            ops = {}
            for s, _, _ in data[st:en + 1]:
                o = s.split()[1] if s.startswith('@') else s.split()[0]
                o = o.split('.')[0]; ops[o] = ops.get(o, 0) + 1
            top = sorted(ops.items(), key=lambda x: -x[1])[:5]
            print(f"{st:5d}-{en:5d} exec={n:9d} len={l:4d} tot={n*l/1e6:7.1f}M ({100*n*l/tot:4.1f}%) samp={sm:5d} ({100*sm/max(ts,1):4.1f}%) {top}")
    if len(sys.argv) > 4:
        x, y = map(int, sys.argv[4].split('-'))
        for i in range(x, y): print(i, data[i][1], data[i][2], data[i][0])
    print("--- top sampled instructions")
    for i in sorted(range(len(data)), key=lambda i: -data[i][2])[:14]:
        print(i, data[i][1], data[i][2], data[i][0])
