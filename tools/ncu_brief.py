"""One paragraph per kernel from an .ncu-rep: time, launch shape, occupancy, issue utilisation, pipes, DRAM, top stalls.
python tools/ncu_brief.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def g(r, k):
    return r[hdr.index(k)] if k in hdr else ''
for r in rows[2:]:
    st = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and r[i] not in ('', '0'):
            st.append((int(float(r[i])), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
    tot = sum(s for s, _ in st) or 1
    print(g(r, 'Kernel Name')[:60])
    print(f"  {g(r,'gpu__time_duration.sum')} us  grid {g(r,'launch__grid_size')} x {g(r,'launch__block_size')}  regs {g(r,'launch__registers_per_thread')}  smem {g(r,'launch__shared_mem_per_block_allocated')} KB  "
          f"warps_active {float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):.1f}%  issue_active {float(g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f}%  inst {float(g(r,'smsp__inst_executed.sum'))/1e6:.1f}M  thr/inst {g(r,'smsp__thread_inst_executed_per_inst_executed.ratio')}")
    print(f"  pipes: alu {float(g(r,'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active')):.0f}% fma {float(g(r,'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')):.0f}% lsu {float(g(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active')):.0f}% xu {float(g(r,'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active')):.0f}%   dram rd {g(r,'dram__bytes_read.sum')} wr {g(r,'dram__bytes_write.sum')} {units[hdr.index('dram__bytes_read.sum')]} ({float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.0f}% of peak)  smem bank conflicts {g(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')}")
    print('  stalls: ' + ', '.join(f'{n} {100*s/tot:.0f}%' for s, n in sorted(st, reverse=True)[:8]))
