#!/usr/bin/env python
"""Single-frame latency of the HOST seams - the calls the two ROS nodes make once per stereo frame - on one B200, with
ordinary (pageable) host buffers like the std::vector storage of ROS messages, wall clock around the blocking call:
  constructor   mod_scene_flow  : disparity x2 + flow in, PointCloud2 payload (32 B/px) out   (two-process topology)
  clusterer     mod_cluster     : PointCloud2 payload in, MovingObjects out
  fused         mod_pipeline    : disparity x2 + flow in, MovingObjects out (cloud stays on the GPU)
and the reference's construct() + dataCB() on one host thread (they are single-threaded) for the same frame.
    python profiles/bench_single_frame.py > profiles/r01_single_frame.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from moving_object_detector_b200 import Context, synth  # noqa: E402
from helpers import seq_inputs  # noqa: E402


def timed(fn, reps):
    fn()
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts)), float(np.min(ts))


for name in ("gazebo_640x480", "zed_1280x720", "zed2k_2208x1242"):
    cfg = synth.CONFIGS[name]
    seq = synth.make_sequence(cfg, 2)
    a = seq_inputs(seq, 0)
    h, w = a["now"].image.shape
    line = {"case": name, "width": w, "height": h, "host_memory": "pageable", "bytes_in_constructor": w * h * 16,
            "bytes_cloud": w * h * 32}
    default_threads = Context.host_copy_threads()
    with Context(w, h, max_batch=1) as ctx:
        cloud = ctx.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])[0]
        kept = np.empty_like(cloud)  # a cloud message buffer kept across frames (what the host classes do)
        for threads in (0, 1, 2, default_threads, 8):
            Context.set_host_copy_threads(threads)
            sf_fresh = timed(lambda: ctx.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"]), 20)
            sf = timed(lambda: ctx.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"], out=kept), 20)
            cl = timed(lambda: ctx.cluster(cloud, want_labels=False), 20)
            pl = timed(lambda: ctx.pipeline(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"],
                                            want_labels=False), 20)
            # a stream: frame ids (mod_set_frame_ids) let the previous disparity stay on the device; the two images
            # alternate as now / prev so that every call keeps the promise "equal ids, equal contents"
            state = {"k": 1}

            def stream_call(fn):
                k = state["k"]
                state["k"] = k + 1
                ctx.set_frame_ids(k + 1, k)
                now, prev = (a["now"], a["prev"]) if k % 2 else (a["prev"], a["now"])
                return fn(now, prev)
            pl_stream = timed(lambda: stream_call(lambda now, prev: ctx.pipeline(
                now, prev, a["flow"], a["cam"], a["prev2now"], a["dt_sec"], want_labels=False)), 20)
            sf_stream = timed(lambda: stream_call(lambda now, prev: ctx.scene_flow(
                now, prev, a["flow"], a["cam"], a["prev2now"], a["dt_sec"], out=kept)), 20)
            line["copy_threads_%d%s" % (threads, " (driver staging)" if threads == 0 else "")] = {
                "mod_scene_flow_ms": sf[0], "mod_scene_flow_fresh_output_ms": sf_fresh[0], "mod_cluster_ms": cl[0],
                "two_node_ms": sf[0] + cl[0], "mod_pipeline_ms": pl[0],
                "stream_mod_scene_flow_ms": sf_stream[0], "stream_mod_pipeline_ms": pl_stream[0]}
        Context.set_host_copy_threads(default_threads)
        line["default_copy_threads"] = default_threads
        n_obj = ctx.pipeline(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"], want_labels=False).n_objects
        line["objects"] = int(n_obj)
    try:
        import ref_py
        if ref_py.available():
            sp, sn, _ = ref_py.stamps_for(seq["dts"][0])
            disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
            kw = dict(f=seq["f"], T=seq["T"], min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"])
            now, prev, flow = (np.ascontiguousarray(x) for x in (disp[1], disp[0], flows[0]))
            t0 = time.perf_counter()
            r = ref_py.scene_flow(now, prev, flow, seq["cam"], seq["transforms"][0], sp, sn, subscribers=1, **kw)
            t1 = time.perf_counter()
            q = ref_py.cluster(r["points32"], subscribers=1)
            t2 = time.perf_counter()
            line["reference_construct_ms"] = (t1 - t0) * 1e3
            line["reference_dataCB_ms"] = (t2 - t1) * 1e3
            line["reference_objects"] = int(q["n_objects"])
    except Exception as e:  # the reference build is optional on the box
        line["reference_error"] = repr(e)
    print(json.dumps(line), flush=True)
