#!/usr/bin/env python
"""Per-source-line totals of an `ncu -i rep --page source --csv --print-source cuda,sass` dump: warp instructions
executed and stall samples per CUDA line, largest first.  Usage: ncu_lines.py dump.csv [top_n]"""
import collections
import csv
import sys


def main(path, top=40):
    files = {}
    cur_file = None
    hdr = None
    inst = collections.Counter()
    samp = collections.Counter()
    text = {}
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1]
            hdr = None
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row
            li, si = hdr.index("Line No"), hdr.index("Source")
            ii, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(row) <= max(ii, sa):
            continue
        if not row[li].strip().isdigit():
            continue  # a SASS line of the combined view (its totals are already in its CUDA line)
        key = (cur_file.split("/")[-1], int(row[li]))
        try:
            inst[key] += int(row[ii] or 0)
            samp[key] += int(row[sa] or 0)
        except ValueError:
            continue
        text[key] = row[si].strip()[:110]
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for key, n in inst.most_common(int(top)):
        print(f"{key[0]}:{key[1]:5d} inst {n:11d} {100.0*n/tot_i:5.1f}%  samples {samp[key]:7d} {100.0*samp[key]/max(tot_s,1):5.1f}%  {text[key]}")


if __name__ == "__main__":
    main(*sys.argv[1:])
