#!/usr/bin/env python
"""Times mod_image_crop_batched_dev (image_crop) on one B200: device-resident batches of camera
frames, CUDA events, inputs + outputs larger than L2.  Algorithmic bytes = crop area x (source bytes/pixel + 3).
    python profiles/bench_image_crop.py > profiles/r01_image_crop.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moving_object_detector_b200 import Context  # noqa: E402

PEAK = 6544.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

for name, w, h, tw, th, enc, bpp, batch in (("zed_720p_bgra8", 1280, 720, 1024, 576, "bgra8", 4, 128),
                                            ("zed_720p_bgr8", 1280, 720, 1024, 576, "bgr8", 3, 128),
                                            ("hd2k_bgra8", 2208, 1242, 1920, 1080, "bgra8", 4, 48),
                                            ("uhd_bgr8", 3840, 2160, 3200, 1800, "bgr8", 3, 16)):
    src = torch.randint(0, 256, (batch, h, w, bpp), dtype=torch.uint8, device="cuda")
    out = torch.empty((batch, th, tw, 3), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    with Context(w, h, max_batch=1) as c:
        run = lambda: c.image_crop_batched_dev(w, h, batch, enc, src, h * w * bpp, w * bpp, tw, th, out, stream=stream)
        for _ in range(5):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        n = 50
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    algo = batch * tw * th * (bpp + 3)
    print(json.dumps({"kernel": "image_crop_kernel", "case": name, "frames": batch, "src": [w, h], "crop": [tw, th],
                      "encoding": enc, "ms": ms, "frames_per_s": batch / ms * 1e3, "algorithmic_bytes": algo,
                      "achieved_gbs": algo / ms / 1e6, "peak_gbs": PEAK, "frac": algo / ms / 1e6 / PEAK}))
