"""GPU tests of the filtered residual gate (sceneflow.cu): the default scene-flow kernel decides
`|flow - static_flow| >= dynamic_flow_diff` (scene_flow_constructor.cpp:191-198) from a float evaluation with an
error bound and re-evaluates undecided pixels exactly.  Outputs must equal the oracle (and the exact kernel) bit for
bit, in particular for flows placed a few ulps -- or a few milli-pixels -- either side of the threshold."""
import dataclasses

import numpy as np
import pytest

from helpers import assert_cloud_bit_exact, oracle_scene_flow, seq_inputs
from moving_object_detector_b200 import (CameraInfo, Context, SceneFlowConstructorConfig, TransformMsg, synth)

pytestmark = pytest.mark.gpu


def _near_threshold_flows(static, thr, rng, offsets):
    """flow = static flow + a vector of length thr + offset (offset drawn per pixel from `offsets`) in a random
    direction: the residual of every pixel with a static flow sits `offset` pixels from the gate."""
    h, w, _ = static.shape
    ang = rng.uniform(0, 2 * np.pi, (h, w))
    off = rng.choice(np.asarray(offsets, np.float64), (h, w))
    r = thr + off
    flow = np.stack([static[..., 0].astype(np.float64) + r * np.cos(ang),
                     static[..., 1].astype(np.float64) + r * np.sin(ang)], axis=-1).astype(np.float32)
    flow[np.isnan(static).any(axis=-1)] = 0.25
    return np.ascontiguousarray(flow)


@pytest.fixture(scope="module")
def gazebo():
    return synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)


def test_filter_decides_almost_everything_and_matches_oracle(oracle, gazebo):
    inp = seq_inputs(gazebo, 0)
    want = oracle_scene_flow(oracle, gazebo, 0)["points32"]
    with Context(640, 480) as c:
        c.filter_fallbacks()
        got, _, _ = c.scene_flow(**inp)
        n_fb = c.filter_fallbacks()
        c.set_static_flow_filter(False)
        exact, _, _ = c.scene_flow(**inp)
        assert c.filter_fallbacks() == 0
    assert_cloud_bit_exact(got, want, "filtered kernel vs oracle")
    assert_cloud_bit_exact(exact, want, "exact kernel vs oracle")
    assert n_fb < 0.002 * 640 * 480, f"{n_fb} pixels fell back to the exact evaluation"


@pytest.mark.parametrize("thr", [1, 5, 40])
def test_filter_at_the_threshold(oracle, gazebo, thr):
    """Residuals within a few float ulps of the gate: all of these must be re-evaluated exactly."""
    rng = np.random.default_rng(77 + thr)
    static = oracle_scene_flow(oracle, gazebo, 0, want_static=True)["static_flow"]
    ulp = np.spacing(np.float32(thr)).astype(np.float64)
    flow = _near_threshold_flows(static, float(thr), rng, [k * ulp for k in (-3, -2, -1, -0.5, 0, 0.5, 1, 2, 3)])
    inp = dict(seq_inputs(gazebo, 0), flow=flow)
    seq = dict(gazebo, flows=flow[None])
    want = oracle_scene_flow(oracle, seq, 0, dynamic_flow_diff=thr, want_class=True)
    with Context(640, 480) as c:
        c.filter_fallbacks()
        got, _, _ = c.scene_flow(**inp, config=SceneFlowConstructorConfig(dynamic_flow_diff=thr))
        n_fb = c.filter_fallbacks()
    assert_cloud_bit_exact(got, want["points32"], "near-threshold flows")
    cls = want["class"]
    assert (cls == 3).sum() > 1000 and (cls == 2).sum() > 1000, "both sides of the gate must occur"
    assert n_fb > 0.5 * ((cls == 2) | (cls == 3)).sum(), "pixels at the gate must take the exact path"


def test_filter_sweep_of_distances_from_the_threshold(oracle, gazebo):
    """Residuals 1e-7 ... 0.5 px either side of the gate: wherever the filter decides, it must agree."""
    rng = np.random.default_rng(4242)
    static = oracle_scene_flow(oracle, gazebo, 0, want_static=True)["static_flow"]
    mags = [10.0 ** e for e in np.arange(-7, 0, 0.5)]
    flow = _near_threshold_flows(static, 5.0, rng, [s * m for m in mags for s in (-1, 1)])
    inp = dict(seq_inputs(gazebo, 0), flow=flow)
    seq = dict(gazebo, flows=flow[None])
    want = oracle_scene_flow(oracle, seq, 0)
    with Context(640, 480) as c:
        c.filter_fallbacks()
        got, _, _ = c.scene_flow(**inp)
        n_fb = c.filter_fallbacks()
    assert_cloud_bit_exact(got, want["points32"], "distance sweep")
    assert 0 < n_fb < 0.8 * 640 * 480


@pytest.mark.parametrize("case", ["far_translation", "scaled_quaternion", "tiny_focal", "huge_focal", "near_plane",
                                  "negative_min_disparity", "infinite_max_disparity"])
def test_filter_exotic_constants(oracle, case):
    """Constants outside the range the bound was derived for switch the filter off (or push pixels to the exact
    path); results stay the oracle's."""
    cfg = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], w=160, h=96, fx=140.0, cx=80.5, cy=48.5, seed=31337,
                              n_invalid_rects=1, n_movers=1)
    seq = dict(synth.make_sequence(cfg, 2))
    tr = np.array(seq["transforms"][0], np.float64)
    cam = list(seq["cam"])
    if case == "far_translation":
        tr[:3] = [1e15, -3e14, 2e13]
    elif case == "scaled_quaternion":
        tr[3:] *= 37.0
    elif case == "tiny_focal":
        cam[0] = cam[1] = 1e-5
    elif case == "huge_focal":
        cam[0] = cam[1] = 1e12
    elif case == "near_plane":
        tr[:3] = [0.0, 0.0, -20.0]   # moves the back wall onto the camera plane: Z ~ 0 after the transform
    elif case == "negative_min_disparity":
        seq["min_disparity"] = -5.0
        seq["disp"] = seq["disp"].numpy().copy() if hasattr(seq["disp"], "numpy") else np.array(seq["disp"])
        seq["disp"][:, 10:30, 10:60] = -2.5
    elif case == "infinite_max_disparity":
        seq["max_disparity"] = float("inf")
    seq["transforms"] = [tr]
    seq["cam"] = tuple(cam)
    inp = seq_inputs(seq, 0)
    want = oracle_scene_flow(oracle, seq, 0)["points32"]
    with Context(160, 96) as c:
        got, _, _ = c.scene_flow(**inp)
    assert_cloud_bit_exact(got, want, case)


def test_filter_equals_exact_kernel_at_scale():
    """Device-resident 720p batch (the bench workload plus a near-threshold copy): filtered and exact kernels must
    write identical clouds and masks."""
    import torch
    cfg = synth.CONFIGS["zed_1280x720"]
    B = 6
    seq = synth.make_sequence(cfg, B + 1, device="cuda:0")
    disp, flows = seq["disp"], seq["flows"]
    W, H = cfg.w, cfg.h
    cam = CameraInfo.pinhole(*seq["cam"])
    transforms = [TransformMsg(t[:3], t[3:]) for t in seq["transforms"]]
    with Context(W, H, max_batch=B) as c:
        st = torch.empty((B, H * W, 2), dtype=torch.float32, device="cuda:0")
        pts = [torch.empty((B, H * W, 8), dtype=torch.float32, device="cuda:0") for _ in range(2)]
        msk = [torch.empty((B, H * W), dtype=torch.uint8, device="cuda:0") for _ in range(2)]
        stream = torch.cuda.current_stream().cuda_stream

        def run(fl, filt, k, static=None):
            pairs = Context.make_pairs([disp[i + 1] for i in range(B)], [disp[i] for i in range(B)],
                                       [fl[i] for i in range(B)], transforms, seq["dts"], seq["f"], seq["T"],
                                       seq["min_disparity"], seq["max_disparity"])
            c.set_static_flow_filter(filt)
            c.scene_flow_batched_dev(W, H, pairs, cam, SceneFlowConstructorConfig(), 0.3, pts[k], msk[k],
                                     static_flow=static, stream=stream)
            torch.cuda.synchronize()

        def same(a, b):
            return torch.equal(a.view(torch.int32), b.view(torch.int32))

        # (1) the bench workload
        run(flows, False, 0, static=st)
        c.filter_fallbacks()
        run(flows, True, 1)
        n1 = c.filter_fallbacks()
        assert same(pts[0], pts[1]) and torch.equal(msk[0], msk[1])
        assert n1 < 0.002 * B * W * H
        # (2) flows rebuilt around the gate: static flow + (5 + offset) px in a random direction, offsets log-spread
        g = torch.Generator(device="cuda:0").manual_seed(99)
        ang = torch.rand((B, H * W), device="cuda:0", generator=g, dtype=torch.float64) * (2 * np.pi)
        mag = 10.0 ** (torch.rand((B, H * W), device="cuda:0", generator=g, dtype=torch.float64) * 8.0 - 8.0)
        sign = torch.where(torch.rand((B, H * W), device="cuda:0", generator=g) < 0.5, -1.0, 1.0)
        r = 5.0 + sign * mag
        near = torch.stack([st[..., 0].double() + r * torch.cos(ang), st[..., 1].double() + r * torch.sin(ang)],
                           dim=-1).float()
        near = torch.nan_to_num(near, nan=0.25).reshape(B, H, W, 2).contiguous()
        run(near, False, 0)
        c.filter_fallbacks()
        run(near, True, 1)
        n2 = c.filter_fallbacks()
        assert same(pts[0], pts[1]) and torch.equal(msk[0], msk[1])
        assert msk[0].sum().item() > 0.1 * B * W * H and (msk[0] == 0).sum().item() > 0.1 * B * W * H
        assert 0 < n2 < B * W * H
