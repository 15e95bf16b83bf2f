"""Summarise an `ncu --page raw --csv` export: one block per captured kernel with the counters the
roofline discussion uses.    python tools/ncu_summary.py gpurun_out/x_raw.csv [more.csv ...]"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
        "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        ]


def main():
    for path in sys.argv[1:]:
        rows = [r for r in csv.reader(open(path, newline="")) if r]
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        names, units = rows[hdr], rows[hdr + 1]
        for r in rows[hdr + 2:]:
            d = dict(zip(names, r))
            print(f"## {d.get('Kernel Name', '?')[:110]}")
            for k in KEYS:
                if k in d:
                    print(f"  {k:85s} {d[k]:>16s} {units[names.index(k)]}")
            extra = [k for k in names if "fmaheavy" in k and k not in KEYS]
            for k in extra:
                print(f"  {k:85s} {d[k]:>16s} {units[names.index(k)]}")


if __name__ == "__main__":
    main()
