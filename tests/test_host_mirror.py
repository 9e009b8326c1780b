"""The C++ host mirror (moving_object_detector_b200/host): SceneFlowConstructorB200::construct and
ClustererB200::dataCB keep the reference's message layouts and degrade rules and reach the GPU only through the
C ABI.  CPU: it builds, links and refuses to construct without a GPU.  GPU: one frame pair driven through the two
classes equals the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "moving_object_detector_b200", "host")
BIN = os.path.join(HOST, "host_selftest")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "moving_object_detector_b200", "csrc"), "-s"])
    subprocess.check_call(["make", "-C", HOST, "-s"])


def test_host_mirror_builds_and_fails_loudly_without_gpu():
    _build()
    out = subprocess.run([BIN, "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no-gpu checks ok" in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert "no usable sm_100a GPU" in out.stdout


def test_host_mirror_links_only_the_c_abi():
    _build()
    out = subprocess.check_output(["nm", "-D", "--undefined-only", os.path.join(ROOT, "moving_object_detector_b200",
                                                                                "libmod_b200_host.so")], text=True)
    undefined = [l.split()[-1] for l in out.splitlines()]
    assert any(s.startswith("mod_scene_flow") for s in undefined) and any(s.startswith("mod_cluster") for s in undefined)
    assert not any(s.startswith(("cuda", "cu", "orc_")) and not s.startswith("cur") for s in undefined), undefined


@pytest.mark.gpu
def test_host_mirror_frame_pair_equals_oracle(oracle, tmp_path):
    from helpers import assert_cloud_bit_exact
    from moving_object_detector_b200 import synth
    _build()
    seq = synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)
    disp, flow = seq["disp"].numpy(), seq["flows"].numpy()[0]
    h, w = disp.shape[1:]
    tr, cam = seq["transforms"][0], seq["cam"]
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("<2i", w, h))
        f.write(struct.pack("<6d", *cam))
        f.write(struct.pack("<4f", seq["f"], seq["T"], seq["min_disparity"], seq["max_disparity"]))
        f.write(struct.pack("<7d", *tr))
        f.write(struct.pack("<4I", 100, 0, 100, 33333333))           # stamps: dt = 0.033333333 s exactly as ros::Duration
        f.write(struct.pack("<2i", 5, 2500))
        f.write(struct.pack("<2d", 0.15, 0.3))
        f.write(struct.pack("<2i", 4, 0))
        f.write(np.ascontiguousarray(disp[0]).tobytes())
        f.write(np.ascontiguousarray(disp[1]).tobytes())
        f.write(np.ascontiguousarray(flow).tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([BIN, str(case), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = open(out, "rb").read()
    cloud = np.frombuffer(raw[:w * h * 32], np.float32).reshape(h, w, 8)
    n = struct.unpack("<i", raw[w * h * 32:w * h * 32 + 4])[0]
    pos = w * h * 32 + 4
    objs = np.frombuffer(raw[pos:pos + n * 80], np.float64).reshape(n, 10)
    pos += n * 80
    image = np.frombuffer(raw[pos:pos + w * h * 3], np.uint8).reshape(h, w, 3)
    pos += w * h * 3
    n_markers = struct.unpack("<i", raw[pos:pos + 4])[0]
    pos += 4
    markers = []
    for _ in range(n_markers):
        mid, mtype, action, npts = struct.unpack("<4i", raw[pos:pos + 16])
        color = struct.unpack("<4f", raw[pos + 16:pos + 32])
        scale = struct.unpack("<3d", raw[pos + 32:pos + 56])
        pts = np.frombuffer(raw[pos + 56:pos + 56 + npts * 24], np.float64).reshape(npts, 3)
        pos += 56 + npts * 24
        markers.append((mid, mtype, action, color, scale, pts))
    assert pos == len(raw)
    dt = 0.0 + 1e-9 * 33333333.0                                      # Duration.toSec(): sec + 1e-9 * nsec
    want = oracle.scene_flow(np.ascontiguousarray(disp[1]), np.ascontiguousarray(disp[0]), np.ascontiguousarray(flow),
                             cam, tr, dt, f=seq["f"], T=seq["T"], min_disparity=seq["min_disparity"],
                             max_disparity=seq["max_disparity"])
    assert_cloud_bit_exact(cloud, want["points32"], "~scene_flow")
    wc = oracle.cluster(want["points32"])
    assert n == wc["n_objects"] >= 1
    for got, exp in zip(objs, wc["objects"]):
        assert int(got[0]) == exp["id"]
        assert np.array_equal(got[1:4].astype(np.float32), exp["center"])      # float -> float64 widening is lossless
        assert np.array_equal(got[7:10].astype(np.float32), exp["bbox"])
    # ~clusters_image (publishClustersImage) and ~clusters (publishClusters / cluster2Marker)
    lut = np.load(os.path.join(ROOT, "tests", "golden", "hsv_colormap_bgr.npy"))
    assert np.array_equal(image, oracle.clusters_image(wc["labels"], n, lut))
    members = oracle.cluster_members(wc["labels"], n)
    assert n_markers == n + 1 and markers[0][2] == 3 and len(markers[0][5]) == 0          # DELETEALL first
    flat = want["points32"].reshape(-1, 8)
    for k, (mid, mtype, action, color, scale, pts) in enumerate(markers[1:]):
        assert (mid, mtype, action) == (k, 8, 0) and scale == (0.05, 0.05, 0.0)           # POINTS, ADD
        b, g, r = (int(v) for v in lut[k * 255 // n])
        assert color == tuple(np.float32(v) for v in (r / 255.0, g / 255.0, b / 255.0, 1.0))
        assert np.array_equal(pts, flat[members[k], :3].astype(np.float64))                # u-outer / v-inner order
