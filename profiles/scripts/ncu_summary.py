#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: key throughput metrics and top stall reasons per launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'sm__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-' * 100)
    for w in want:
        if w in idx:
            print('%-72s %s %s' % (w, r[idx[w]], units[idx[w]]))
    st = [(h, r[i]) for h, i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    def f(x):
        try:
            return float(x.replace(',', ''))
        except ValueError:
            return 0.0
    print('  top stalls (warps stalled per issue-active cycle):')
    for h, v in sorted(st, key=lambda x: -f(x[1]))[:8]:
        print('    %-70s %s' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
