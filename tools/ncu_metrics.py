#!/usr/bin/env python
"""Print the headline metrics of every kernel in an .ncu-rep: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_metrics.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
want += [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            v = r[i]
            try:
                if 'stalled' in w and float(v) < 1.0:
                    continue
            except ValueError:
                pass
            print(f"{w:<90}{v:>22} {units[i]}")
    print('---')
