"""Host seams with pageable message buffers staged by the library's own copy threads (csrc/hostcopy.cu): results must
not depend on the thread count, the buffer kind (pageable / mod_host_alloc / mod_host_register) or the frame size
(sizes that are not a multiple of the staging chunk, sizes below the staging threshold)."""
import ctypes as C

import numpy as np
import pytest

from helpers import seq_inputs

pytestmark = pytest.mark.gpu


def _run_all(ctx, a, out=None):
    cloud, depth, static = ctx.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"],
                                          want_depth=True, want_static_flow=True, out=out)
    cl = ctx.cluster(cloud, want_labels=True)
    pl = ctx.pipeline(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"], want_points=True,
                      want_labels=True)
    return dict(cloud=cloud[..., [0, 1, 2, 4, 5, 6]].tobytes(), depth=depth.tobytes(), static=static.tobytes(),
                labels=cl.labels.tobytes(), objs=cl.objects.tobytes(), p_cloud=pl.points32[..., [0, 1, 2, 4, 5, 6]].tobytes(),
                p_labels=pl.labels.tobytes(), p_objs=pl.objects.tobytes())


@pytest.mark.parametrize("size", [(640, 480), (1280, 720), (322, 203), (96, 64)])
def test_results_do_not_depend_on_the_copy_threads(mod, size):
    import dataclasses
    from moving_object_detector_b200 import Context, synth
    w, h = size
    base = synth.CONFIGS["gazebo_640x480"]
    cfg = dataclasses.replace(base, w=w, h=h, fx=base.fx * w / base.w, cx=base.cx * w / base.w, cy=base.cy * h / base.h)
    a = seq_inputs(synth.make_sequence(cfg, 2), 0)
    before = Context.host_copy_threads()
    try:
        with Context(w, h) as ctx:
            Context.set_host_copy_threads(0)   # plain cudaMemcpyAsync: the driver stages
            assert Context.host_copy_threads() == 0
            want = _run_all(ctx, a)
            for n in (1, 2, 4, 7):
                Context.set_host_copy_threads(n)
                assert Context.host_copy_threads() == n
                got = _run_all(ctx, a)
                for k in want:
                    assert got[k] == want[k], (n, k)
    finally:
        Context.set_host_copy_threads(before)


def test_pinned_and_registered_buffers_take_the_direct_path(mod):
    from moving_object_detector_b200 import Context, DisparityImage, synth
    lib = mod.load()
    a = seq_inputs(synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2), 0)
    h, w = a["now"].image.shape
    with Context(w, h) as ctx:
        want = _run_all(ctx, a)
        # output buffer from mod_host_alloc, inputs registered in place
        nbytes = w * h * 32
        p = lib.mod_host_alloc(nbytes)
        assert p
        try:
            out = np.frombuffer((C.c_uint8 * nbytes).from_address(p), np.float32).reshape(h, w, 8)
            flow = np.ascontiguousarray(a["flow"])
            assert lib.mod_host_register(flow.ctypes.data, flow.nbytes) == 0
            try:
                b = dict(a, flow=flow)
                got = _run_all(ctx, b, out=out)
            finally:
                assert lib.mod_host_unregister(flow.ctypes.data) == 0
            for k in want:
                assert got[k] == want[k], k
        finally:
            lib.mod_host_free(p)
    assert lib.mod_host_register(None, 16) != 0 and lib.mod_host_unregister(None) != 0


def test_sequence_host_call_gives_the_same_objects_from_pageable_and_pinned_frames(mod):
    """mod_pipeline_sequence_host stages pageable frames itself and sends pinned ones with one DMA each: same results,
    with the copy threads on and off."""
    import dataclasses
    import torch
    from moving_object_detector_b200 import (CameraInfo, ClustererConfig, Context, OBJECT_DTYPE, STATS_DTYPE,
                                             SceneFlowConstructorConfig, TransformMsg, synth)
    cfg = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], seed=2103)
    n_pairs, max_objs = 5, 64
    seq = synth.make_sequence(cfg, n_pairs + 1)
    disp, flows = seq["disp"].contiguous(), seq["flows"].contiguous()
    H, W = disp.shape[1:]
    before = Context.host_copy_threads()

    def run(d, f):
        objs = np.zeros((n_pairs, max_objs), OBJECT_DTYPE)
        stats = np.zeros(n_pairs, STATS_DTYPE)
        with Context(W, H, max_batch=2) as c:
            c.pipeline_sequence_host(d, f, CameraInfo.pinhole(*seq["cam"]),
                                     [TransformMsg(t[:3], t[3:]) for t in seq["transforms"]], seq["dts"], seq["f"],
                                     seq["T"], seq["min_disparity"], seq["max_disparity"], SceneFlowConstructorConfig(),
                                     ClustererConfig(), objs, stats, max_objs)
        return objs.tobytes(), stats.tobytes()

    try:
        Context.set_host_copy_threads(0)
        want = run(disp.numpy(), flows.numpy())
        assert np.frombuffer(want[1], STATS_DTYPE)["n_objects"].sum() > 0
        for threads in (1, 4):
            Context.set_host_copy_threads(threads)
            assert run(disp.numpy(), flows.numpy()) == want, threads
        assert run(disp.pin_memory(), flows.pin_memory()) == want
    finally:
        Context.set_host_copy_threads(before)


def test_previous_disparity_stays_resident_with_frame_ids(mod):
    """mod_set_frame_ids: in a stream the image uploaded as `now` is reused as the next call's `prev`.  Same results as
    without ids; the reuse happens (resident_hits); a wrong promise is not made by the library itself: without ids,
    or after a size change, everything is uploaded."""
    import dataclasses
    from moving_object_detector_b200 import Context, DisparityImage, synth
    cfg = dataclasses.replace(synth.CONFIGS["gazebo_640x480"], seed=2104)
    n = 5
    seq = synth.make_sequence(cfg, n + 1)
    H, W = seq["disp"].shape[1:]

    def run(use_ids, fused):
        outs = []
        with Context(W, H) as c:
            for i in range(n):
                a = seq_inputs(seq, i)
                if use_ids:
                    c.set_frame_ids(1000 + i + 1, 1000 + i)
                if fused:
                    r = c.pipeline(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"], want_points=True)
                    outs.append((r.points32[..., [0, 1, 2, 4, 5, 6]].tobytes(), r.labels.tobytes(), r.objects.tobytes()))
                else:
                    cloud = c.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])[0]
                    outs.append(cloud[..., [0, 1, 2, 4, 5, 6]].tobytes())
            return outs, c.resident_hits

    for fused in (False, True):
        want, hits0 = run(False, fused)
        got, hits = run(True, fused)
        assert hits0 == 0 and hits == n - 1
        assert got == want

    # ids that do not chain (prev_id is not the previous now_id), and a call without ids in between: no reuse
    with Context(W, H) as c:
        for i in range(3):
            a = seq_inputs(seq, i)
            c.set_frame_ids(50 + 2 * i, 7)
            c.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])
        a = seq_inputs(seq, 3)
        c.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])   # no ids: forgets
        c.set_frame_ids(60, 54)
        c.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])
        assert c.resident_hits == 0


def test_cluster_resident_equals_the_round_trip(mod):
    """mod_cluster_resident on the cloud mod_scene_flow left on the device == mod_cluster on the downloaded cloud;
    refused when there is no such cloud."""
    from moving_object_detector_b200 import ClustererConfig, Context, ModError, synth
    a = seq_inputs(synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2), 0)
    h, w = a["now"].image.shape
    with Context(w, h) as c:
        with pytest.raises(ModError):
            c.cluster_resident(w, h)
        cloud = c.scene_flow(a["now"], a["prev"], a["flow"], a["cam"], a["prev2now"], a["dt_sec"])[0]
        got = c.cluster_resident(w, h, ClustererConfig())
        want = c.cluster(cloud, ClustererConfig())
        assert got.n_objects == want.n_objects >= 2
        assert got.labels.tobytes() == want.labels.tobytes() and got.objects.tobytes() == want.objects.tobytes()
        with pytest.raises(ModError):   # mod_cluster replaced the resident cloud by the caller's
            c.cluster_resident(w, h)
