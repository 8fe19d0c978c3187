"""Print the handful of raw ncu metrics used in profiles/*.md from a .ncu-rep (last kernel in the report)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]; v = r[-1]
keys = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.per_cycle_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__block_size",
        "launch__grid_size", "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum")
for k, x in zip(h, v):
    if k in keys: print(k, x)
    if "issue_stalled" in k and "per_issue_active" in k and x not in ("0", "0.000000"): print("  stall", k.split("stalled_")[1].split("_per")[0], x)
