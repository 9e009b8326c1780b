#!/usr/bin/env python
"""Clusterer stress (BASELINE.json configs[3]): 2208x1242 frames with dense moving-object clusters - a frame-spanning
serpentine, blobs straddling cluster_size, thousands of specks, checkerboards, depth-stepped blobs (synth.make_stress_cloud)
- through mod_cluster_batched_dev on one B200, device-resident, CUDA events, for neighbor_distance 1 / 4 / 10; the
reference's own clusterer (oracle/_ref/libmod_ref.so, ClustererNodelet::dataCB) timed beside it on one host thread
(the reference's callback is single-threaded) on one frame.
    python profiles/bench_clusterer_stress.py > profiles/r01_clusterer_stress.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from moving_object_detector_b200 import ClustererConfig, Context, OBJECT_DTYPE, STATS_DTYPE, synth  # noqa: E402

W, H, FRAMES, MAX_OBJS = 2208, 1242, 8, 1024
SPEED = 0.3

clouds = np.stack([synth.make_stress_cloud(W, H, seed=4004 + i)[0] for i in range(FRAMES)])
v = clouds[..., 4:7]
dyn = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]) >= SPEED
points = torch.from_numpy(clouds).cuda()
objs = torch.zeros(FRAMES * MAX_OBJS * OBJECT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
stats = torch.zeros(FRAMES * STATS_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream

try:
    import ref_py
    have_ref = ref_py.available()
except Exception:
    have_ref = False

with Context(W, H, max_batch=FRAMES) as ctx:
    for n in (1, 4, 10):
        cfg = ClustererConfig(cluster_size=2500, depth_diff=0.15, dynamic_speed=SPEED, neighbor_distance=n)
        # mask = None: the dynamic map is computed from the cloud (calculateDynamicMap), as for a foreign PointCloud2
        run = lambda: ctx.cluster_batched_dev(W, H, FRAMES, points, None, cfg, objs, stats, MAX_OBJS, stream=stream)
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        st = np.frombuffer(stats.cpu().numpy().tobytes(), STATS_DTYPE)
        line = {"case": "zed2k_2208x1242 clusterer stress", "neighbor_distance": n, "frames": FRAMES,
                "dynamic_pixels": int(dyn.sum()), "dynamic_fraction": float(dyn.mean()), "ms_per_batch": ms,
                "ms_per_frame": ms / FRAMES, "frames_per_s": FRAMES / ms * 1e3,
                "mpix_per_s": FRAMES * W * H / ms / 1e3, "objects": int(st["n_objects"].sum()) if "n_objects" in st.dtype.names else None,
                "status_errors": int((st["status"] != 0).sum()) if "status" in st.dtype.names else None}
        if have_ref:
            t0 = time.perf_counter()
            r = ref_py.cluster(clouds[0], cluster_size=2500, depth_diff=0.15, dynamic_speed=SPEED, neighbor_distance=n,
                               subscribers=1)
            cpu_ms = (time.perf_counter() - t0) * 1e3
            line["reference_cpu_ms_per_frame_1thread"] = cpu_ms
            line["reference_objects_frame0"] = int(r["n_objects"])
        print(json.dumps(line), flush=True)
