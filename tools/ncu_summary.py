"""Condense `ncu --page details --csv` / `--page raw --csv` exports (made on the GPU box, see tools/_run.sh history in
DESIGN.md section 6) into the per-kernel text summaries committed under profiles/.

    python tools/ncu_summary.py <label> <details.csv> <raw.csv> [...more triples]  > profiles/rNN_xxx.txt
"""
import csv
import sys

DETAILS = ["Duration", "SM Frequency", "Elapsed Cycles", "Compute (SM) Throughput", "Memory Throughput", "DRAM Throughput",
           "L1/TEX Cache Throughput", "L2 Cache Throughput", "Executed Ipc Active", "Issue Slots Busy", "Mem Busy", "Mem Pipes Busy",
           "Registers Per Thread", "Dynamic Shared Memory Per Block", "Achieved Occupancy", "No Eligible", "Eligible Warps Per Scheduler",
           "Issued Warp Per Scheduler", "Warp Cycles Per Issued Instruction", "Block Size", "Grid Size"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
       "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
       "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def rows(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def main():
    args = sys.argv[1:]
    for i in range(0, len(args), 3):
        label, det, raw = args[i:i + 3]
        d = rows(det)
        print("=" * 110)
        print(label)
        print("kernel:", d[0]["Kernel Name"][:160] if d else "?")
        for key in DETAILS:
            for r in d:
                if r["Metric Name"] == key:
                    print("  %-44s %14s %s" % (key, r["Metric Value"], r["Metric Unit"]))
                    break
        r = rows(raw)
        if len(r) >= 2:                       # first row = units, second = values
            units, vals = r[0], r[1]
            print("  -- raw metrics")
            for key in RAW:
                if key in vals:
                    print("  %-92s %16s %s" % (key, vals[key], units.get(key, "")))
            try:
                mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = sum(float(vals[k].replace(",", "")) * mul[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                print("  %-92s %16.1f Mbyte" % ("dram__bytes_read.sum + dram__bytes_write.sum", tot / 1e6))
            except (KeyError, ValueError):
                pass


if __name__ == "__main__":
    main()
