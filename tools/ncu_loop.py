"""Hot loop of a kernel from `ncu --page source --csv`: the SASS lines executed at least `frac` x the most executed one,
in program order.  python tools/ncu_loop.py file.csv [frac]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]
idx = {h: i for i, h in enumerate(hdr)}
seen, d2 = set(), []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[idx['Instructions Executed']].isdigit():
        continue
    if r[idx['Address']] in seen:
        break
    seen.add(r[idx['Address']])
    d2.append(r)
tot = sum(int(r[idx['Instructions Executed']]) for r in d2)
ts = sum(int(r[idx['# Samples']]) for r in d2)
mx = max(int(r[idx['Instructions Executed']]) for r in d2)
print('SASS lines', len(d2), 'warp instructions', tot, 'samples', ts, 'max exec', mx)
n = e_sum = s_sum = 0
for r in d2:
    e = int(r[idx['Instructions Executed']])
    if e > frac * mx:
        n += 1
        e_sum += e
        s_sum += int(r[idx['# Samples']])
        print(f"{e:9d} {r[idx['# Samples']]:>6} {r[idx['Avg. Threads Executed']]:>5} {r[idx['Source']].strip()[:100]}")
print('hot lines', n, 'share of instructions %.1f%%' % (100 * e_sum / tot), 'share of samples %.1f%%' % (100 * s_sum / ts))
