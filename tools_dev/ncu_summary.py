"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one block per kernel) into a short text file."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__cluster_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg.per_second',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
for d in rows[2:]:
    print("-" * 100)
    for i, k in enumerate(hdr):
        if i < len(d) and (k in KEYS or 'pipe_tensor' in k or 'tmem' in k.lower()) and d[i] not in ('', 'n/a'):
            print(f"{k:96s} {d[i][:80]} {units[i]}")
