#!/usr/bin/env python
"""Summarises an `ncu --csv` launch list (gpu__time_duration etc.) per kernel.  Usage: summarize_launches.py file.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui = (hdr.index(c) for c in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit'))
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].split('::')[-1]
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
        agg.setdefault(name, collections.OrderedDict()).setdefault(r[mi], []).append(v * scale)
    total = 0.0
    print(f"{'kernel':28s} {'n':>3s} {'time_us':>10s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s} {'warp_inst':>14s} "
          f"{'l2_atom_sect':>13s} {'l2_red_sect':>12s}")
    for k, m in agg.items():
        mean = lambda key: (sum(m[key]) / len(m[key])) if key in m else float('nan')
        t = mean('gpu__time_duration.sum')
        total += t
        print(f"{k:28s} {len(m['gpu__time_duration.sum']):3d} {t:10.1f} {mean('dram__bytes_read.sum'):11.1f} "
              f"{mean('dram__bytes_write.sum'):11.1f} {mean('smsp__inst_executed.sum'):14.0f} "
              f"{mean('lts__t_sectors_op_atom.sum'):13.0f} {mean('lts__t_sectors_op_red.sum'):12.0f}")
    print(f"sum of per-kernel means: {total:.1f} us")


if __name__ == '__main__':
    main(sys.argv[1])
