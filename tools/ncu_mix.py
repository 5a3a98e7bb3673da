#!/usr/bin/env python3
"""Instruction mix of one kernel from `ncu --page source --csv`: executed warp instructions by SASS opcode."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
seen = set(); mix = collections.Counter(); samp = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[idx['# Samples']].isdigit() or r[idx['Address']] in seen:
        continue
    seen.add(r[idx['Address']])
    toks = r[idx['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
    op = '.'.join(op.split('.')[:2]) if len(sys.argv) > 2 else op.split('.')[0]
    mix[op] += int(r[idx['Instructions Executed']]); samp[op] += int(r[idx['# Samples']])
tot = sum(mix.values()); ts = sum(samp.values())
print(f"total warp instructions {tot}, samples {ts}")
for op, n in mix.most_common(40):
    print(f"{op:14s} {n:12d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(ts,1):5.1f}%")
