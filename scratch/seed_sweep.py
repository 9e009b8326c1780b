#!/usr/bin/env python
"""Step time of the benchmark workload for the eight per-rank seeds of a weak-scaling run, on ONE GPU: how much of the
N = 8 step time is the spread of the data (rank r runs seed 2002 + r) and not the system.
    python scratch/seed_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
for off in range(8):
    wl = bench.Workload("zed_1280x720", 63, dev, 0, seed_offset=off)
    for _ in range(5):
        wl.step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(200):
        wl.step()
    b.record()
    torch.cuda.synchronize()
    print("seed offset %d: %.4f ms per step" % (off, a.elapsed_time(b) / 200), flush=True)
    wl.close()
    del wl
    torch.cuda.empty_cache()
