#!/usr/bin/env python
"""cuobjdump -sass of the built library, per kernel: instruction count and the mnemonics that prove which hardware
path a kernel uses (tcgen05: UTCHMMA / UTCQMMA, tensor-memory loads LDTM / stores STTM, TMA UTMALDG / UTMASTG,
cp.async LDGSTS, warp reductions REDUX / CREDUX, legacy tensor cores HMMA).  usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "marllb_b200/libmarllb_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "LDGSTS", "REDUX", "CREDUX", "HMMA", "FFMA", "DFMA", "STL", "LDL"]
blocks = out.split("Function : ")[1:]
print(f"# cuobjdump -sass {lib}: static instruction counts per kernel (sm_100a)")
print("kernel,instructions," + ",".join(KEYS))
for blk, name in zip(blocks, names):
    ops = collections.Counter()
    n = 0
    for line in blk.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            n += 1
            ops[m.group(1)] += 1
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "")
    print(f"\"{short[:90]}\",{n}," + ",".join(str(ops.get(k, 0)) for k in KEYS))
