"""Print selected metrics from an `ncu --page raw --csv` dump (one column per metric)."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'sm__throughput.avg.pct', 'launch__registers_per_thread', 'sm__warps_active.avg.pct', 'launch__occupancy_limit',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'launch__shared_mem_per_block', 'launch__grid_size', 'launch__block_size', 'launch__waves',
        'sm__inst_executed_pipe_', 'smsp__inst_executed_pipe_', 'local', 'issue_stalled', 'lts__t_sector_hit_rate',
        'l1tex__t_sector_hit_rate', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg ', 'smsp__thread_inst_executed_per_inst',
        'smsp__inst_executed_per_warp', 'achieved_occupancy', 'lts__t_bytes.sum ', 'l1tex__t_bytes.sum ']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print('=' * 100)
    for i, h in enumerate(hdr):
        if h in ('Kernel Name', 'ID') or any(k in h for k in KEYS):
            v = row[i]
            if v in ('', '0', 'n/a') and h not in ('Kernel Name',):
                continue
            print(f"{h:90s} {units[i]:14s} {v}")
