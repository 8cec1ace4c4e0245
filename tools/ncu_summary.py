#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, without a GPU) into a small text table for profiles/: per captured launch the
duration, DRAM/L2/L1 traffic, issue-slot utilisation, SIMT efficiency (threads per issued instruction), pipe
utilisation and the top stall reasons.  usage: ncu_summary.py <file.ncu-rep> [out.txt]"""
import csv
import io
import subprocess
import sys

M = [("gpu__time_duration.sum", "duration"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
     ("smsp__inst_executed.sum", "warp instructions"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
     ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction (SIMT efficiency, of 32)"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
     ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
     ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
     ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
     ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
     ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
     ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
     ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
     ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
     ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction")]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# %s" % rep]
    for k, r in enumerate(rows[2:]):
        out.append("\n## launch %d: %s" % (k, r[idx["Kernel Name"]]))
        for m, label in M:
            if m in idx:
                out.append("%-52s %18s %s" % (label, r[idx[m]], units[idx[m]]))
    txt = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(txt)
    else:
        print(txt)


if __name__ == "__main__":
    main()
