#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: executed warp-instructions by opcode and the
hottest address ranges.  Usage: ncu_source_hot.py src.csv [n_elements]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
nelem = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ia, isrc, iex, ism = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samples = Counter(), Counter()
tot = 0
data = []
for r in rows[2:]:
    try:
        ex = int(r[iex])
    except ValueError:
        continue
    src = r[isrc].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] += ex
    samples[op] += int(r[ism] or 0)
    tot += ex
    data.append((r[ia], ex, int(r[ism] or 0), src))
print("total warp-instructions executed:", tot)
if nelem:
    print("thread-instructions per element (32 x warp-instr / elements): %.1f" % (tot * 32 / nelem))
print("%-10s %14s %7s %9s" % ("opcode", "warp-instr", "share", "samples"))
for op, n in ops.most_common(28):
    extra = "  %.2f/elem" % (n * 32 / nelem) if nelem else ""
    print("%-10s %14d %6.1f%% %9d%s" % (op, n, 100.0 * n / tot, samples[op], extra))
if "--dump" in sys.argv:
    for a, ex, sm, src in data:
        print(a, ex, sm, src)
