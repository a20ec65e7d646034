"""Thread- and warp-level instruction totals of one ncu capture:  python tools/ncu_inst.py file.ncu-rep
(sass__thread_inst_executed_true_per_opcode summed, smsp__inst_executed.sum, FP32 / FP64 thread instructions by opcode)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[2]
get = lambda name: next((vals[i] for i, h in enumerate(hdr) if h == name), None)
print("kernel", get("Kernel Name"))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sass__thread_inst_executed_true_per_opcode", "smsp__thread_inst_executed_per_inst_executed.ratio",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
          "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"):
    print(f"{k:75s} {get(k)}")
