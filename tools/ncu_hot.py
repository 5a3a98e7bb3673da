"""Top stall-sample instructions from `ncu --page source --csv` output: python tools/ncu_hot.py file.csv [min_share]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
seen = set()
data = []
for r in rows[hi + 1:]:
    if len(r) == len(hdr) and r[idx['# Samples']].isdigit() and r[idx['Address']] not in seen:
        seen.add(r[idx['Address']])
        data.append(r)
tot = sum(int(r[idx['# Samples']]) for r in data)
texec = sum(int(r[idx['Instructions Executed']]) for r in data)
print('total samples', tot, 'warp instructions executed', texec, 'SASS lines', len(data))
print(f"{'samples':>8} {'share':>6} {'exec':>10} {'thr/inst':>8}  source")
for r in data:
    s = int(r[idx['# Samples']])
    if s > tot * thr:
        print(f"{s:8d} {s / tot * 100:5.1f}% {int(r[idx['Instructions Executed']]):10d} {r[idx['Avg. Threads Executed']]:>8}  {r[idx['Source']].strip()[:100]}")
