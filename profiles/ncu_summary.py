#!/usr/bin/env python
''' Print the metrics of interest from an `ncu --page raw --csv` dump:  python profiles/ncu_summary.py raw.csv [kernel-substring] '''
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ''
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__waves_per_multiprocessor', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if flt not in name:
        continue
    print('---', name[:60], 'id', r[hdr.index('ID')])
    for w in want:
        if w in hdr:
            print(f'  {w:80s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}')
