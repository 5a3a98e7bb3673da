#!/bin/bash
# usage (on the GPU box): tools/ab.sh name [ENV=val ...] [-- bench args]: runs bench.py and prints a one-line summary
name=$1; shift
envs=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
[ "$1" == "--" ] && shift
env "${envs[@]}" python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
python - "$name" <<PY
import json,sys
f=sys.argv[1]
try:
    d=json.load(open("gpurun_out/bench_%s.json"%f))
    print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"]*1e3) for k,v in d["kernels"].items()}, "rounds", d["decode_stats"]["sync_rounds"], "launches", d["gpu_launches"])
except Exception as e:
    print(f, "ERR", e); print(open("gpurun_out/bench_%s.err"%f).read()[-2000:])
PY
