"""Print selected metrics per profiled launch from `ncu -i X.ncu-rep --page raw --csv > X.csv`."""
import csv, sys
WANT = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1])))
sel = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 else None
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
if sel:
    data = [data[i] for i in sel]
for w in WANT:
    if w in idx:
        print(w[-62:].ljust(64), rows[1][idx[w]][:8].ljust(9), " | ".join(r[idx[w]][:12].rjust(12) for r in data))
    else:
        print("MISSING", w)
