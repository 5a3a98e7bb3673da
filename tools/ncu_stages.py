#!/usr/bin/env python3
"""Per-stage breakdown of one kernel from `ncu --page source --csv`: the SASS is cut at every BAR.SYNC; for each
stage prints executed warp instructions per warp, the share of stall samples and the opcode mix.
usage: ncu_stages.py source.csv warps_launched"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2])
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
seen = set(); region = 0; tot = collections.Counter(); samp = collections.Counter(); mix = collections.defaultdict(collections.Counter)
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[idx['# Samples']].isdigit() or r[idx['Address']] in seen:
        continue
    seen.add(r[idx['Address']])
    src = r[idx['Source']]; ex = int(r[idx['Instructions Executed']]); sm = int(r[idx['# Samples']])
    tot[region] += ex; samp[region] += sm
    toks = src.split(); op = toks[1] if toks[0].startswith('@') else toks[0]
    mix[region][op.split('.')[0]] += ex
    if 'BAR.SYNC' in src:
        region += 1
print(f"total {sum(tot.values()) / warps:.0f} warp instructions per warp, {sum(samp.values())} samples")
for i in range(region + 1):
    print(i, f"instr/warp {tot[i] / warps:.0f}", f"samples {100 * samp[i] / max(1, sum(samp.values())):.1f}%",
          ' '.join(f"{k}:{v / warps:.0f}" for k, v in mix[i].most_common(16)))
