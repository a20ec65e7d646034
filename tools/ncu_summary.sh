#!/bin/bash
# usage: ncusum.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','l1tex__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__inst_executed.sum','smsp__inst_executed.avg.per_cycle_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','launch__grid_size','launch__block_size','sm__cycles_elapsed.avg','smsp__sass_thread_inst_executed_op_ffma_pred_on.sum','smsp__sass_thread_inst_executed_op_fmul_pred_on.sum','smsp__sass_thread_inst_executed_op_fadd_pred_on.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(f'{w:75s} {vals[i]:>18s} {units[i]}')
print('--- stall reasons (warp issue stalled, pct of samples)')
st=[(float(vals[i].replace(',','')),h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') or (h.startswith('smsp__average_warp_latency_issue_stalled') )]
for v,h in sorted(st,reverse=True)[:10]: print(f'{h:90s} {v:10.3f}')
"
