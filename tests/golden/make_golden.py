#!/usr/bin/env python
"""Generates the committed golden fixtures tests/golden/*.npz FROM THE REFERENCE'S OWN CODE.

The reference ships no golden vectors.  Its hot-path translation units are compiled here, unmodified, from
/root/reference against stand-in dependency headers (oracle/Makefile target ref_full -> oracle/_ref/libmod_ref.so,
bound by oracle/ref_py.py); this script runs SceneFlowConstructor::construct() and ClustererNodelet::dataCB() on
seeded inputs and freezes what the two nodes publish.  While generating, the CPU oracle (oracle/libmod_oracle.so) is
required to agree bit for bit; only its per-pixel class image, which the reference does not expose, is taken from
the oracle.  tests/test_golden.py then checks that (1) the oracle, (2) the independent numpy restatement
(tests/np_restatement.py) and (3, -m gpu) the CUDA path reproduce the fixtures; tests/test_reference_build.py
checks that the reference build still does.  Needs /root/reference (this container only):

    make -C oracle && python tests/golden/make_golden.py
"""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import oracle_py  # noqa: E402
import ref_py  # noqa: E402
from moving_object_detector_b200 import synth  # noqa: E402


def scene_flow_case(name, cfg, pair=0, **kw):
    seq = synth.make_sequence(cfg, pair + 2)
    disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
    now, prev, flow = disp[pair + 1].copy(), disp[pair].copy(), flows[pair].copy()
    for key, fn in kw.items():
        if key == "mutate":
            fn(now, prev, flow)
    tr = np.asarray(seq["transforms"][pair], np.float64)
    # ROS stamps have nanosecond resolution: dt is what (stamp_now - stamp_prev).toSec() returns for them
    stamp_prev, stamp_now, dt = ref_py.stamps_for(float(seq["dts"][pair]))
    cam = np.asarray(seq["cam"], np.float64)
    meta = np.asarray([seq["f"], seq["T"], seq["min_disparity"], seq["max_disparity"]], np.float32)
    want = oracle_py.scene_flow(now, prev, flow, cam, tr, dt, dynamic_flow_diff=5, f=float(meta[0]), T=float(meta[1]),
                                min_disparity=float(meta[2]), max_disparity=float(meta[3]), want_depth=True,
                                want_static=True, want_class=True)
    ref = ref_py.scene_flow(now, prev, flow, cam, tr, stamp_prev, stamp_now, dynamic_flow_diff=5, f=float(meta[0]),
                            T=float(meta[1]), min_disparity=float(meta[2]), max_disparity=float(meta[3]))
    assert ref["status"] == 0 and ref["published"] == 7
    fields = [0, 1, 2, 4, 5, 6]
    for a, b in ((ref["points32"][..., fields], want["points32"][..., fields]), (ref["depth"], want["depth"]),
                 (ref["static_flow"], want["static_flow"])):
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)].view(np.uint32),
                                                                           b[~np.isnan(b)].view(np.uint32)), name
    pts = ref["points32"].copy()
    pts[..., 3] = 0.0   # the record's pad floats are uninitialised upstream
    pts[..., 7] = 0.0
    want = dict(want, points32=pts, depth=ref["depth"], static_flow=ref["static_flow"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), disp_now=now, disp_prev=prev, flow=flow, cam=cam,
                        transform=tr, dt=np.float64(dt), meta=meta, dynamic_flow_diff=np.int32(5),
                        points32=want["points32"], depth=want["depth"], static_flow=want["static_flow"],
                        cls=want["class"])
    counts = np.bincount(want["class"].ravel(), minlength=4)
    print(f"{name}: {now.shape[1]}x{now.shape[0]} classes I..IV = {counts.tolist()}")
    return want["points32"]


def cluster_case(name, cloud, **params):
    want = oracle_py.cluster(cloud, **params)
    ref = ref_py.cluster(cloud, **params)
    assert ref["status"] == 0 and ref["objects_lossless"]
    assert np.array_equal(ref["labels"], want["labels"]) and np.array_equal(ref["dynamic"], want["dynamic"]), name
    objects = ref["objects"].copy()
    objects["n_points"] = [int((ref["labels"] == k).sum()) for k in objects["id"]]   # not in the message
    assert objects.tobytes() == want["objects"].tobytes(), name
    want = dict(want, labels=ref["labels"], dynamic=ref["dynamic"], objects=objects)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), points32=cloud,
                        params=np.asarray([params["cluster_size"], params["neighbor_distance"]], np.int32),
                        thresholds=np.asarray([params["depth_diff"], params["dynamic_speed"]], np.float64),
                        labels=want["labels"], dynamic=want["dynamic"], objects=want["objects"])
    print(f"{name}: {want['n_objects']} objects, {int((want['labels'] >= 0).sum())} labelled pixels")


def main():
    small = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], w=96, h=72, fx=83.0, cx=48.5, cy=36.5, seed=7001,
                                n_invalid_rects=2)
    cloud = scene_flow_case("sf_small_ego", small)
    ident = dataclasses.replace(small, identity_ego=True, seed=7002)
    scene_flow_case("sf_small_identity", ident)

    def edge_values(now, prev, flow):
        # every invalid encoding on both frames, flows that round exactly at .5, huge / inf flow, negative zero
        now[0, :8] = [0.0, -0.0, -1.0, np.nan, np.inf, -np.inf, 200.0, 1e-30]
        prev[1, :8] = [0.0, -0.0, -1.0, np.nan, np.inf, -np.inf, 200.0, 1e-30]
        flow[2, :6, 0] = [0.5, -0.5, 1.5, 2.5, np.inf, -np.inf]
        flow[3, :4, 1] = [0.5, 1e30, -1e30, np.nan]
    scene_flow_case("sf_small_edges", dataclasses.replace(small, seed=7003), mutate=edge_values)

    # clusterer: the scene-flow cloud above (real pipeline output) at three window sizes, and a synthetic stress cloud
    for n in (1, 4, 10):
        cluster_case(f"cl_small_n{n}", cloud, cluster_size=20, depth_diff=0.15, dynamic_speed=0.3,
                     neighbor_distance=n)
    stress, _ = synth.make_stress_cloud(224, 160, seed=7004, cluster_size=100, fx=100.0)
    for n in (1, 4, 10):
        cluster_case(f"cl_stress_n{n}", stress, cluster_size=100, depth_diff=0.15, dynamic_speed=0.3,
                     neighbor_distance=n)


if __name__ == "__main__":
    main()
