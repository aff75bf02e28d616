#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump: per-kernel headline metrics, stall reasons, instruction mix."""
import csv
import sys

path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
HEAD = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'smsp__inst_executed_op_global_st.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'sm__cycles_elapsed.avg']
stall = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h]
for d in data:
    name = d[idx['Kernel Name']]
    if filt and filt not in name:
        continue
    print('====', name[:90])
    for k in HEAD:
        if k in idx:
            print('  %-75s %s %s' % (k, d[idx[k]], units[idx[k]]))
    vals = sorted([(float(d[idx[k]] or 0), k) for k in stall], reverse=True)[:8]
    print('  stalls (warps per issue):', ', '.join('%s=%.2f' % (k.split('issue_stalled_')[1].replace('.ratio', '').replace('_per_warp_active', ''), v) for v, k in vals))
