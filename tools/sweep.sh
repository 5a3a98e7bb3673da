#!/bin/bash
# sweep of the entropy-decode tuning knobs on the bench workload (run on the GPU box)
for sb in 256 512 1024 2048; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --skip-pixel-check --sub-bits $sb --relay-rounds 10 > gpurun_out/sweep_sb$sb.json 2>gpurun_out/sweep_sb$sb.err || tail -3 gpurun_out/sweep_sb$sb.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/sweep_sb*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f,'bad'); continue
    k=d['kernels']
    print(f, 'value %.0f'%d['value'], 'ms %.3f'%d['ms_per_step'], ' '.join(f"{n}={v['ms_per_step']*1000:.0f}" for n,v in k.items()), d['decode_stats']['sync_rounds'])
PY
