#!/bin/bash
# usage (GPU box): tools/whatif.sh kernel_regex variant...: duration of one launch of the kernel per variant library (ncu,
# gpu__time_duration only).  For timing experiments with variants whose RESULTS may be wrong: the bench's parity gate
# stops the run after the launch has been measured.
k=$1; shift
for v in "$@"; do
  lib=variants/$v/libkpeg_cuda.so; [ "$v" == "default" ] && lib=libkpeg_b200/lib/libkpeg_cuda.so
  KPEG_BENCH_WHATIF=1 KPEG_CUDA_LIB=$lib ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$k" -c 1 --csv --log-file gpurun_out/whatif_$v.csv \
     python bench.py --steps 1 --warmup 0 --no-cpu-baseline --skip-pixel-check --no-extras --quick > gpurun_out/whatif_$v.log 2>&1
  echo "$v: $(grep gpu__time_duration gpurun_out/whatif_$v.csv | awk -F, '{print $(NF)}' | tr -d '"' | tr '\n' ' ')"
done
