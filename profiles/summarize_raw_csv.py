"""Summarise an `ncu --page raw --csv` export (one row per profiled launch): python profiles/summarize_raw_csv.py raw.csv"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__inst_executed.sum", "inst"), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conflicts"),
        ("launch__occupancy_limit_registers", "occ_regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("sm__cycles_elapsed.avg.per_second", "clk"),
        ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_longsb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_shortsb"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_branch"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "st_sleep"),
        ("smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio", "st_tex"),
        ("smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "st_drain"),
        ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "st_imc"),
        ("smsp__average_warps_issue_stalled_selected_per_issue_active.ratio", "st_selected")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("#%s %s" % (r[hdr.index("ID")], name[:150]))
        out = []
        for key, short in WANT:
            cols = [i for i, h in enumerate(hdr) if h == key or h.endswith("." + key)]
            if cols:
                out.append("%s=%s%s" % (short, r[cols[0]], (" " + units[cols[0]]) if units[cols[0]] and short in ("time", "dram_rd", "dram_wr") else ""))
        print("   " + "  ".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
