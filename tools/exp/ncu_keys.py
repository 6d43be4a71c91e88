"""Print selected raw metrics of one or more ncu reports side by side.  usage: ncu_keys.py rep1 [rep2 ...]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "memory_l1_wavefronts_shared_ideal", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
cols = []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; r = rows[2]
    ix = {k: i for i, k in enumerate(h)}
    cols.append({k: r[ix[k]] if k in ix else "-" for k in KEYS})
for k in KEYS:
    print(k.ljust(92), *[c[k].rjust(16) for c in cols])
