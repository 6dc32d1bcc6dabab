"""Key metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv) as a markdown table.  usage: ncu_summary.py file.ncu-rep [title]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__shared_mem_per_block_static", "static smem / block"),
    ("sm__warps_active.avg.per_cycle_active", "resident warps / SM (active cycles)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction"),
    ("smsp__issue_active.avg.per_cycle_active", "issue slots busy / cycle"),
    ("sm__inst_executed_pipe_fp64.sum", "FP64 warp instructions"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy % (active)"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe busy % (elapsed)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall/issue: wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall/issue: short scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall/issue: long scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall/issue: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall/issue: branch resolving"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall/issue: no instruction"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall/issue: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall/issue: lg throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall/issue: mio throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall/issue: dispatch"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall/issue: not selected"),
]
ix = {h: i for i, h in enumerate(hdr)}
names = [r[ix["Kernel Name"]][:60] for r in data]
print(f"# {title}\n")
print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
print("|---|" + "---:|" * len(names))
for key, label in WANT:
    if key not in ix:
        continue
    vals = [r[ix[key]] for r in data]
    print(f"| {label} (`{key}`, {units[ix[key]]}) | " + " | ".join(vals) + " |")
