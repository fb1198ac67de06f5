"""Key metrics per kernel from an `ncu --page raw --csv` dump (one row per launch)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_issue_stalled_short_scoreboard_per_warp_active.pct", "sm__sass_inst_executed_op_shared_st.sum"]
idx = {}
for w in want:
    for i, h in enumerate(hdr):
        if h == w or h.endswith(w):
            idx[w] = i
            break
for r in rows[2:]:
    print("-----")
    for w in want:
        if w in idx:
            print(f"  {w:90s} {r[idx[w]]:>18s} {units[idx[w]]}")
