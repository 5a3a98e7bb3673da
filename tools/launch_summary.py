"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
   python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            k = d['Kernel Name'].split('(')[0][-48:]
            v = float(d['Metric Value'].replace(',', ''))
            if d['Metric Unit'] in ('us', 'usecond'):
                v *= 1e3
            elif d['Metric Unit'] in ('ms', 'msecond'):
                v *= 1e6
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':50s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
for k, a in agg.items():
    print(f"{k:50s} {a[0]:8d} {a[1] / 1e3:10.1f} {a[1] / a[0] / 1e3:9.2f} {a[1] / tot * 100:6.1f}%")
print(f"{'TOTAL':50s} {sum(a[0] for a in agg.values()):8d} {tot / 1e3:10.1f}")
