#!/usr/bin/env python
"""Turn a .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__sass_thread_inst_executed_op_fp64_pred_on.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % path)
    for r in rows[2:]:
        print("\n## launch id %s: %s  grid %s block %s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")],
                                                          r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        for w in WANT:
            if w in hdr:
                print("%-82s %16s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))


if __name__ == "__main__":
    main(sys.argv[1])
