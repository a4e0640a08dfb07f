#!/usr/bin/env python
"""Key metrics + stall breakdown per kernel from `ncu -i rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_uniform.sum']
stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            print(f"  {w:80s} {r[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h.split('issue_stalled_')[1].split('_per_')[0]) for h in stall), reverse=True)
    print("  stalls/issue: " + ", ".join(f"{n} {v:.2f}" for v, n in st if v >= 0.05))
