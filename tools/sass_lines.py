#!/usr/bin/env python3
"""Static SASS instruction count per CUDA source line for one kernel.
usage: sass_lines.py <nvdisasm -g -c output> <kernel name substring> [top N]"""
import re, sys, collections
txt = open(sys.argv[1]).read().splitlines()
name = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
i0 = [i for i, l in enumerate(txt) if l.lstrip().startswith('.section') and '.text.' in l and name in l][0]
i1 = next((i for i in range(i0 + 1, len(txt)) if txt[i].lstrip().startswith('.section')), len(txt))
cnt = collections.Counter(); cur = None; total = 0
fl = re.compile(r'//## File "([^"]+)", line (\d+)')
for l in txt[i0:i1]:
    m = fl.search(l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        cnt[cur] += 1; total += 1
src = {}
def line(f, n):
    if f not in src:
        try: src[f] = open(f).read().splitlines()
        except Exception: src[f] = []
    return src[f][n - 1].strip()[:100] if 0 < n <= len(src[f]) else ''
print('total static instructions', total)
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    if k: print(f"{v:4d} {k[0].split('/')[-1]}:{k[1]}  {line(*k)}")
