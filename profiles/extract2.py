#!/usr/bin/env python
"""ncu -i <rep> --page raw --csv -> compact per-launch CSV with the metrics profiles/README.md cites (round 2 set: adds the
L1 data-pipe wavefronts, issue activity and the dominant stall reasons).  python profiles/extract2.py rep.ncu-rep > out.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    w = csv.writer(sys.stdout)
    w.writerow([k + (" [" + units[i] + "]" if units[i] else "") for k, i in idx])
    for d in data:
        w.writerow([d[i][:60] for _, i in idx])


if __name__ == "__main__":
    main(sys.argv[1])
