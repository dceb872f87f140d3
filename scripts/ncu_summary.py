"""Summarise an .ncu-rep (raw page) into one line per launch: python scripts/ncu_summary.py rep [out.csv]"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_registers", "occ_reg"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("smsp__inst_executed.sum", "inst"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio")]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
names = [n for k, n in want if k in col]
print(",".join(names), file=out)
for r in rows[2:]:
    vals = []
    for k, n in want:
        if k not in col: continue
        v = r[col[k]]
        if n == "kernel": v = v.split("(")[0].replace(",", ";")[:60]
        else: v = v.replace(",", "") + ("" if n in ("regs", "grid", "block", "inst", "bank_conf", "smem_wavefronts") else " " + units[col[k]])
        vals.append(v)
    print(",".join(vals), file=out)
