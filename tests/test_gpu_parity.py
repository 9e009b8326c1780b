"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): xyz, dynamic mask, labels, ids, bbox, centre bit-exact; velocity within 1e-4
relative -- in fact asserted bit-exact here, because the kernels reproduce the reference's rounding sequence."""
import dataclasses

import numpy as np
import pytest

from helpers import (assert_cloud_bit_exact, assert_objects_match, assert_same_partition, oracle_scene_flow,
                     seq_inputs)
from moving_object_detector_b200 import ClustererConfig, Context, SceneFlowConstructorConfig, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = Context(1280, 720, max_batch=8)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gazebo_seq():
    return synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 3)


@pytest.mark.parametrize("pair", [0, 1])
def test_scene_flow_bit_exact_gazebo(ctx, oracle, gazebo_seq, pair):
    inp = seq_inputs(gazebo_seq, pair)
    got, depth, static = ctx.scene_flow(**inp, want_depth=True, want_static_flow=True)
    want = oracle_scene_flow(oracle, gazebo_seq, pair, want_depth=True, want_static=True, want_class=True)
    assert_cloud_bit_exact(got, want["points32"], "scene flow cloud")
    assert_cloud_bit_exact(depth[..., None], want["depth"][..., None], "depth")
    assert_cloud_bit_exact(static, want["static_flow"], "static flow")
    counts = np.bincount(want["class"].ravel(), minlength=4)
    assert (counts > 0).all(), f"synthetic data must exercise all four pixel classes: {counts}"


def test_scene_flow_identity_ego(ctx, oracle):
    cfg = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], identity_ego=True, seed=1002)
    seq = synth.make_sequence(cfg, 2)
    got, _, _ = ctx.scene_flow(**seq_inputs(seq, 0))
    want = oracle_scene_flow(oracle, seq, 0)
    assert_cloud_bit_exact(got, want["points32"])


@pytest.mark.parametrize("thr", [1, 5, 100])
def test_scene_flow_threshold_values(ctx, oracle, gazebo_seq, thr):
    got, _, _ = ctx.scene_flow(**seq_inputs(gazebo_seq, 0), config=SceneFlowConstructorConfig(dynamic_flow_diff=thr))
    want = oracle_scene_flow(oracle, gazebo_seq, 0, dynamic_flow_diff=thr)
    assert_cloud_bit_exact(got, want["points32"])


def test_cluster_matches_oracle_gazebo(ctx, oracle, gazebo_seq):
    cloud = oracle_scene_flow(oracle, gazebo_seq, 0)["points32"]
    want = oracle.cluster(cloud)
    got = ctx.cluster(cloud, ClustererConfig())
    assert want["n_objects"] >= 2
    assert np.array_equal(got.labels, want["labels"]), "labels (ids in the reference's order) differ"
    assert_objects_match(got.objects, want["objects"], cloud, want["labels"])


def test_pipeline_fused_matches_oracle(ctx, oracle, gazebo_seq):
    inp = seq_inputs(gazebo_seq, 1)
    res = ctx.pipeline(**inp, want_points=True, want_labels=True)
    want_sf = oracle_scene_flow(oracle, gazebo_seq, 1)
    assert_cloud_bit_exact(res.points32, want_sf["points32"])
    want = oracle.cluster(want_sf["points32"])
    assert np.array_equal(res.labels, want["labels"])
    assert_objects_match(res.objects, want["objects"], want_sf["points32"], want["labels"])


def test_cluster_long_root_chains_are_repeatable(oracle):
    """Regression: a frame-spanning serpentine at N=1 builds long chains of tile roots; a path-halving store in
    cl_roots used to race with the final-root store and misroute a few member pixels (flaky median).  Fresh contexts
    and repeated calls must give the oracle's objects every time."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cl_stress_n1.npz"))
    cloud = np.ascontiguousarray(g["points32"])
    cfg = ClustererConfig(int(g["params"][0]), float(g["thresholds"][0]), float(g["thresholds"][1]), 1)
    for trial in range(25):
        with Context(256, 192, max_batch=1) as c:
            for rep in range(2):
                got = c.cluster(cloud, cfg)
                assert np.array_equal(got.labels, g["labels"]), f"trial {trial} rep {rep}"
                assert_objects_match(got.objects, g["objects"], cloud, g["labels"])


@pytest.mark.parametrize("w,h", [(100, 37), (98, 41), (97, 33), (64, 8), (68, 9), (2, 2), (1, 5), (260, 16)])
def test_scene_flow_every_kernel_path_and_ragged_sizes(oracle, w, h):
    """Even widths take the two-pixel-per-thread kernel (partial tiles on both edges), odd widths the one-pixel
    kernel; with the static-flow image requested the exact kernel runs, without it the filtered one.  All must equal
    the oracle bit for bit (the prefetch of rows below the image / beyond the batch must stay in bounds)."""
    cfg = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], w=w, h=h, fx=80.0, cx=w / 2 + 0.5, cy=h / 2 + 0.5,
                              seed=9000 + w, n_invalid_rects=1, n_movers=1)
    seq = synth.make_sequence(cfg, 2)
    with Context(max(w, 8), max(h, 8), max_batch=1) as c:
        got, depth, static = c.scene_flow(**seq_inputs(seq, 0), want_depth=True, want_static_flow=True)
        got_f, depth_f, _ = c.scene_flow(**seq_inputs(seq, 0), want_depth=True)
    want = oracle_scene_flow(oracle, seq, 0, want_depth=True, want_static=True)
    assert_cloud_bit_exact(got, want["points32"], "cloud")
    assert_cloud_bit_exact(got_f, want["points32"], "cloud (filtered kernel)")
    assert_cloud_bit_exact(depth_f[..., None], want["depth"][..., None], "depth (filtered kernel)")
    assert_cloud_bit_exact(depth[..., None], want["depth"][..., None], "depth")
    assert_cloud_bit_exact(static, want["static_flow"], "static flow")
