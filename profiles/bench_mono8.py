#!/usr/bin/env python
"""MONO8 conversion (mod_image_to_mono8_batched_dev) on one B200: 512 cropped 1024x576 frames per launch (both cameras
of 256 stereo pairs; smaller batches time the host-side launch path instead of the kernel), device-resident, CUDA events.  Algorithmic bytes = pixels x (source bytes per pixel + 1).
    python profiles/bench_mono8.py > profiles/r02_mono8.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moving_object_detector_b200 import Context  # noqa: E402

W, H, BATCH = 1024, 576, 512
peak = 6544.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
stream = torch.cuda.current_stream().cuda_stream
with Context(W, H, max_batch=1) as ctx:
    for enc, bpp in (("bgr8", 3), ("bgra8", 4), ("mono8", 1)):
        src = torch.randint(0, 256, (BATCH, H, W * bpp), dtype=torch.uint8, device="cuda")
        out = torch.empty((BATCH, H, W), dtype=torch.uint8, device="cuda")
        run = lambda: ctx.to_mono8_batched_dev(W, H, BATCH, enc, src, H * W * bpp, W * bpp, 0, out, stream=stream)
        for _ in range(3):
            run()
        # 20 launches back to back between two events (a single launch behind an idle GPU would time the host's launch
        # latency); source + result of one launch (>= 151 MB) exceed the 126 MB L2, so every launch streams from HBM
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps
        nbytes = BATCH * H * W * (bpp + 1)
        print(json.dumps({"case": f"to_mono8 {enc} {BATCH}x{W}x{H}", "ms": t, "gbs": nbytes / t / 1e6,
                          "frac_of_measured_hbm_peak": nbytes / t / 1e6 / peak, "peak_gbs": peak,
                          "frames_per_s": BATCH / t * 1e3}), flush=True)
