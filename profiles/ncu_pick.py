#!/usr/bin/env python
"""Prints selected metrics of an `ncu -i rep --page raw --csv` dump, one block per profiled launch.
Usage: ncu -i x.ncu-rep --page raw --csv | python profiles/ncu_pick.py [substring ...]"""
import csv
import sys

DEFAULT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit", "smsp__issue_active.avg.pct",
           "issue_stalled", "bank_conflicts", "thread_inst_executed_per_inst_executed",
           "sm__inst_executed_pipe_", "smsp__inst_executed_op_", "sm__throughput.avg.pct", "l1tex__throughput.avg.pct",
           "lts__throughput.avg.pct", "dram__throughput.avg.pct", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]


def main():
    pats = sys.argv[1:] or DEFAULT
    rows = list(csv.reader(sys.stdin))
    hdr = rows[0]
    for r in rows[2:]:
        print("-----")
        for i, h in enumerate(hdr):
            if any(p in h for p in pats) and i < len(r) and r[i] not in ("", "n/a"):
                print(f"{h} = {r[i]} {rows[1][i]}")


if __name__ == "__main__":
    main()
