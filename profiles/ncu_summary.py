'''
Turn an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into the per-kernel summary kept under profiles/: duration, DRAM bytes,
throughputs, occupancy, registers, instruction count and the top warp-stall reasons of every captured launch.
    ncu -i gpurun_out/fused.ncu-rep --page raw --csv | python profiles/ncu_summary.py > profiles/r2/fused_ncu_full.txt
'''
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__cycles_active.avg', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
units = rows[1]
stall = [i for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
name_i = hdr.index('Kernel Name')
for r in rows[2:]:
    print(f'--- {r[name_i][:90]}  id {r[0]}')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'  {k:<82} {r[i]:>16} {units[i]}')
    st = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, hdr[i]) for i in stall), reverse=True)[:5]
    for v, h in st:
        print(f'  {h:<82} {v:>16.3f}')
