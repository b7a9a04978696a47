"""Key metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv), one row per launch."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__inst_executed.sum", "winst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"), ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"), ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"), ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"), ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_disp")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(r[idx["Kernel Name"]].split("(")[0])
        for key, name in WANT:
            if key in idx:
                print(f"    {name:10s} {r[idx[key]]:>16s} {units[idx[key]]}")


if __name__ == "__main__":
    main(sys.argv[1])
