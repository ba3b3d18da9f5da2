"""Selected metrics of every kernel in an `ncu --set full` report, as CSV (what profiles/*_ncu_*.csv hold):
  python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep > profiles/<name>.csv"""
import csv, io, subprocess, sys
METRICS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",   # tensor MATH active
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",      # tensor-core unit busy (math + operand fetch)
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",  # tensor-core shared-memory operand reads
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size",
]
txt = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(txt)))
h, units = rows[0], rows[1]
cols = [m for m in METRICS if m in h]
w = csv.writer(sys.stdout)
w.writerow(["Kernel Name"] + cols)
w.writerow([""] + [units[h.index(m)] for m in cols])
for r in rows[2:]:
    d = dict(zip(h, r))
    w.writerow([d["Kernel Name"][:110]] + [d[m] for m in cols])
