#!/usr/bin/env python
"""Per-kernel summary of an .ncu-rep (ncu -i ... --page raw --csv): duration, DRAM bytes, throughputs, stalls."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("sm__inst_executed.sum", "inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_nsel"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for row in rows[2:]:
        name = row[idx["Kernel Name"]].split("(")[0]
        print(f"== {name}  grid={row[idx['Grid Size']]} block={row[idx['Block Size']]}")
        parts = []
        for key, short in KEYS:
            if key in idx:
                parts.append(f"{short}={row[idx[key]]}{units[idx[key]] if short in ('dur','dram_rd','dram_wr') else ''}")
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
