#!/usr/bin/env python
"""Summarise the SASS page of an ncu report: runs of instructions with equal execution counts."""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 3e6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Instructions Executed" in r)
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows:
    try:
        data.append((r[1].strip(), int(r[iI]), int(r[iS])))
    except Exception:
        pass
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print("total warp instr", tot, "samples", ts, "static instrs", len(data))
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
        print(f"{st:5d}-{en:5d} exec={n:9d} len={l:4d} tot={n*l/1e6:7.1f}M ({100*n*l/tot:4.1f}%) samp={sm:5d} ({100*sm/ts:4.1f}%) {top}")
if len(sys.argv) > 3:
    a, b = map(int, sys.argv[3].split('-'))
    for i in range(a, b): print(i, data[i][1], data[i][2], data[i][0])
print("--- top sampled instructions")
for i in sorted(range(len(data)), key=lambda i: -data[i][2])[:25]:
    print(i, data[i][1], data[i][2], data[i][0])
