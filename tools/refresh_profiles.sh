#!/bin/bash
# usage (here, after tools/gpu_round.sh <tag> ran on the GPU box): tools/refresh_profiles.sh <tag> -- copies the bench lines and
# regenerates the ncu / SASS summaries under profiles/ from gpurun_out/<tag>_*
set -e
cd "$(dirname "$0")/.."
T=$1
cp gpurun_out/${T}_bench.json profiles/r2_bench.json
cp gpurun_out/${T}_bench_reference_arm.json profiles/r2_bench_reference_arm.json
{ echo "# round 2, final state: ncu launch list of 'python bench.py --steps 2 --warmup 1 --no-cpu-baseline --skip-pixel-check --no-extras --quick'"
  echo "# (--metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: the SHARE of the step is what to compare."
  echo "# The list mixes the 8-image device-resident jobs with the 1-image chunks of the end-to-end leg, hence the small averages.)"
  python tools/launch_summary.py gpurun_out/${T}_launches.csv; } > profiles/r2_launches.txt
{ echo "# round 2, final state: ncu --set full --clock-control none, one 8-image job (8 x 3840x2160 q95) per kernel; tools/ncu_brief.py"
  python tools/ncu_brief.py gpurun_out/${T}_prof.ncu-rep; } > profiles/r2_ncu_kernels.txt
for k in entropy_relay_full entropy_cold idct_kernel; do ncu -i gpurun_out/${T}_prof.ncu-rep --page source --csv --kernel-name regex:$k > /tmp/src_$k.csv 2>/dev/null; done
{ echo "# round 2: the symbol loop of entropy_relay_full_kernel (emitting pass), SASS lines executed at least half as often as the most executed one, in program order"
  echo "# columns: warp-level executions, stall samples, average active threads, instruction   (tools/ncu_loop.py on ncu --page source)"
  python tools/ncu_loop.py /tmp/src_entropy_relay_full.csv 0.5; echo
  echo "# the same for entropy_cold_kernel (no records, no DC sums)"; python tools/ncu_loop.py /tmp/src_entropy_cold.csv 0.5; } > profiles/r2_hot_entropy_symbol_loops.txt
{ echo "# round 2: idct_kernel<3>, SASS cut at every BAR.SYNC: 0 = wait for the bulk copies, 1 = pull the block into registers, 2 = dequantise + transform + tie test + sample store,"
  echo "# 3 = colour + pixel stores, 4 = tie records, 5.. = rare exact paths (tools/ncu_stages.py; 97200 warps)"
  python tools/ncu_stages.py /tmp/src_idct_kernel.csv 97200; echo; python tools/ncu_hot.py /tmp/src_idct_kernel.csv 0.015; } > profiles/r2_hot_idct_kernel.txt
echo refreshed profiles from $T
