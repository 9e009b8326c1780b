#!/usr/bin/env python
"""Generates the committed golden fixtures tests/golden/*.npz.

The reference ships no golden vectors and cannot be built or imported in this image, so these
fixtures are produced by the CPU oracle (oracle/libmod_oracle.so) and frozen here; tests/test_golden.py checks that
(1) the oracle still reproduces them, (2) the independent numpy restatement (tests/np_restatement.py) reproduces
them and (3, -m gpu) the CUDA path reproduces them.  Re-run only when the oracle is deliberately changed:

    python tests/golden/make_golden.py
"""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import oracle_py  # noqa: E402
from moving_object_detector_b200 import synth  # noqa: E402


def scene_flow_case(name, cfg, pair=0, **kw):
    seq = synth.make_sequence(cfg, pair + 2)
    disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
    now, prev, flow = disp[pair + 1].copy(), disp[pair].copy(), flows[pair].copy()
    for key, fn in kw.items():
        if key == "mutate":
            fn(now, prev, flow)
    tr, dt = np.asarray(seq["transforms"][pair], np.float64), float(seq["dts"][pair])
    cam = np.asarray(seq["cam"], np.float64)
    meta = np.asarray([seq["f"], seq["T"], seq["min_disparity"], seq["max_disparity"]], np.float32)
    want = oracle_py.scene_flow(now, prev, flow, cam, tr, dt, dynamic_flow_diff=5, f=float(meta[0]), T=float(meta[1]),
                                min_disparity=float(meta[2]), max_disparity=float(meta[3]), want_depth=True,
                                want_static=True, want_class=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), disp_now=now, disp_prev=prev, flow=flow, cam=cam,
                        transform=tr, dt=np.float64(dt), meta=meta, dynamic_flow_diff=np.int32(5),
                        points32=want["points32"], depth=want["depth"], static_flow=want["static_flow"],
                        cls=want["class"])
    counts = np.bincount(want["class"].ravel(), minlength=4)
    print(f"{name}: {now.shape[1]}x{now.shape[0]} classes I..IV = {counts.tolist()}")
    return want["points32"]


def cluster_case(name, cloud, **params):
    want = oracle_py.cluster(cloud, **params)
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
