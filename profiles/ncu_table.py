"""One line per kernel of an .ncu-rep: time, DRAM traffic, warp instructions, issue / occupancy / lanes, top stall reasons.
usage: ncu_table.py <rep> [<rep> ...]"""
import csv
import subprocess
import sys

M = {"t_us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum", "winst": "smsp__inst_executed.sum",
     "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "bank": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
     "regs": "launch__registers_per_thread", "l2hit": "lts__t_sector_hit_rate.pct",
     "long": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "short": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
     "bar": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
     "wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "br": "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
     "nsel": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"}


def scale(v, unit):
    v = float(v.replace(",", "")) if v not in ("", "n/a") else 0.0
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}.get(unit, 1)


for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}")
    print(f"{'kernel':24s} {'us':>8s} {'rdMB':>7s} {'wrMB':>6s} {'Mwinst':>7s} {'lanes':>5s} {'occ%':>5s} {'iss%':>5s} {'Mbank':>6s} {'regs':>4s} {'l2hit':>5s} | stalls/issue: long short bar mio wait br nsel lg")
    for r in rows[2:]:
        g = lambda k: scale(r[idx[M[k]]], units[idx[M[k]]]) if M[k] in idx else 0.0
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        print(f"{name:24s} {g('t_us'):8.1f} {g('rd')/1e6:7.1f} {g('wr')/1e6:6.1f} {g('winst')/1e6:7.1f} {g('lanes'):5.1f} {g('occ'):5.1f} {g('issue'):5.1f} {g('bank')/1e6:6.2f} {g('regs'):4.0f} {g('l2hit'):5.1f} | "
              f"{g('long'):.2f} {g('short'):.2f} {g('bar'):.2f} {g('mio'):.2f} {g('wait'):.2f} {g('br'):.2f} {g('nsel'):.2f} {g('lg'):.2f}")
