"""`ncu -i prof.ncu-rep --page raw --csv` -> JSON with the metrics the design notes cite, one entry per captured launch."""
import csv, json, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__cluster_dim_x", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_no_instructions",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_sleeping",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
        # fp64 kernels of the solve
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "smsp__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__shared_mem_per_block_dynamic", "launch__block_size"]
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    e = {"kernel": r[col["Kernel Name"]][:120]}
    for k in KEEP:
        if k in col:
            e[k] = f"{r[col[k]]} {units[col[k]]}".strip()
    out.append(e)
json.dump({"command": " ".join(sys.argv[2:]), "kernels": out}, sys.stdout, indent=1)
