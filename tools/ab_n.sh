#!/bin/bash
# usage (GPU box with N GPUs): tools/ab_n.sh N name [ENV=val ...]: bench.py under torchrun on N GPUs (no extra legs), one-line summary
N=$1; name=$2; shift 2
env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extras --quick > gpurun_out/benchn_$name.json 2> gpurun_out/benchn_$name.err
python - "$name" <<PY
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/benchn_%s.json"%f).read().strip().splitlines()[-1])
    print(f, "N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["e2e"]["ms_per_step"],2), "d2h GB/s per GPU", round(d["e2e"]["pcie_d2h_gbs_measured"],1))
except Exception as e:
    print(f, "ERR", e); print(open("gpurun_out/benchn_%s.err"%f).read()[-1500:])
PY
