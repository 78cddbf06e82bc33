#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into a compact per-kernel table: python tools/ncu_summary.py rep [out.csv]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur"), ("launch__grid_size", "grid"), ("launch__block_size", "blk"),
        ("launch__registers_per_thread", "regs"), ("launch__waves_per_multiprocessor", "waves"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__inst_executed.sum", "winst"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel")]
idx = [(n, h.index(k)) for k, n in want if k in h]
out = [[n for n, _ in idx]]
for r in rows[2:]:
    line = []
    for n, i in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void ", "")[:44]
        else:
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}" + (units[i][0] if n in ("dur", "dramR", "dramW") and units[i] else "")
            except ValueError:
                pass
        line.append(v)
    out.append(line)
# transpose print
for j in range(len(out[0])):
    print(f"{out[0][j]:10s}", *[f"{o[j]:>22s}" for o in out[1:]])
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as fh:
        csv.writer(fh).writerows(out)
