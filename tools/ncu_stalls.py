#!/usr/bin/env python3
"""Stall-reason totals of one kernel from `ncu --page source --csv`, and the SASS lines that hold most of one reason.
usage: ncu_stalls.py source.csv [reason]   (reason e.g. stall_long_sb)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
seen = set(); data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[idx['# Samples']].isdigit() or r[idx['Address']] in seen:
        continue
    seen.add(r[idx['Address']]); data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
print('total samples', tot)
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k in sorted(keys, key=lambda k: -sum(int(r[idx[k]] or 0) for r in data)):
    s = sum(int(r[idx[k]] or 0) for r in data)
    if s:
        print(f"{k:26s} {s:8d} {100 * s / tot:5.1f}%")
why = sys.argv[2] if len(sys.argv) > 2 else 'stall_long_sb'
print('lines with the most', why)
for r in sorted(data, key=lambda r: -int(r[idx[why]] or 0))[:10]:
    print(f"{r[idx[why]]:>7} {r[idx['Instructions Executed']]:>10}  {r[idx['Source']][:100]}")
