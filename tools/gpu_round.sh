#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh tag -- GPU tests, the full bench line, the reference arm, the ncu launch list and one
# full capture of the wide kernels (each ncu pass only after its command has exited 0 without ncu)
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.txt
tail -3 gpurun_out/${tag}_tests.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -20 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err || tail -5 gpurun_out/${tag}_bench_reference_arm.err
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/%s_bench.json"%sys.argv[1]).read().strip().splitlines()[-1])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["ms_per_step"]*1e3) for k,v in d["kernels"].items()}, d["decode_stats"], d["clocks"])
except Exception as e:
    print("bench ERR", e)
PY
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip-pixel-check --no-extras --quick"
$B > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu1.log 2>&1
$B > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'idct_kernel|idct_patch|expand_kernel|entropy_relay_full|entropy_cold|unstuff_count|unstuff_write|entropy_relay_loop' -s 24 -c 8 -o gpurun_out/${tag}_prof -f $B > gpurun_out/${tag}_ncu2.log 2>&1
tail -3 gpurun_out/${tag}_ncu2.log
