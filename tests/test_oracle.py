"""CPU checks of the oracle itself (test infrastructure): against the one reference translation unit that builds
here (oracle/_ref, the label-equivalence table), against the independent numpy restatement on randomised and
adversarial inputs, and against size-independent properties of the path."""
import os

import numpy as np
import pytest

import np_restatement as npr
from helpers import assert_cloud_bit_exact


# ---- label-equivalence table vs the REFERENCE's own lookup_table.cpp ----------------------------------------------
def _random_ops(rng, n_labels, n_ops):
    ops = [(0, 0, 0)] * n_labels
    for _ in range(n_ops):
        if rng.rand() < 0.6:
            ops.append((1, rng.randint(n_labels), rng.randint(n_labels)))
        else:
            ops.append((2, rng.randint(n_labels), 0))
    ops += [(2, i, 0) for i in range(n_labels)]
    return np.asarray(ops, np.int32)


@pytest.mark.parametrize("seed", range(8))
def test_equivalence_table_matches_reference_lookup_table(oracle, seed):
    if not os.path.exists(oracle.REF_LIB_PATH):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    rng = np.random.RandomState(seed)
    n = int(rng.choice([1, 2, 17, 300, 5000]))
    ops = _random_ops(rng, n, 4 * n)
    ours = oracle.lookup_table_run(n + 8, ops)
    ref = oracle.lookup_table_run(n + 8, ops, ref=True)
    assert np.array_equal(ours, ref)


# ---- scene flow: oracle vs numpy restatement on randomised inputs --------------------------------------------------
def _random_pair(rng, w, h, exotic):
    z = rng.uniform(0.5, 30.0, (h, w))
    f, T = np.float32(rng.uniform(50, 900)), np.float32(rng.uniform(0.05, 0.6))
    disp = (float(f) * float(T) / z + rng.normal(0, 0.3, (h, w))).astype(np.float32)
    prev = (disp + rng.normal(0, 0.5, (h, w))).astype(np.float32)
    flow = rng.normal(0, 3.0, (h, w, 2)).astype(np.float32)
    if exotic:
        enc = np.array([0.0, -0.0, -1.0, np.nan, np.inf, -np.inf, 500.0, 1e-38, 1e-44, 3e38], np.float32)
        for img in (disp, prev):
            m = rng.rand(h, w) < 0.15
            img[m] = enc[rng.randint(len(enc), size=int(m.sum()))]
        fenc = np.array([0.5, -0.5, 1.5, -2.5, np.nan, np.inf, -np.inf, 1e30, -1e30, 2147483648.0, 0.49999997],
                        np.float32)
        m = rng.rand(h, w, 2) < 0.1
        flow[m] = fenc[rng.randint(len(fenc), size=int(m.sum()))]
    q = rng.normal(0, 1, 4)
    q = q / np.linalg.norm(q) if not exotic else q * rng.uniform(0.5, 1.5)   # the reference never normalises
    if rng.rand() < 0.5:
        q = np.array([0.0, np.sin(0.004), 0.0, np.cos(0.004)])               # viso2-like small yaw
    tr = np.concatenate([rng.normal(0, 0.05, 3), q])
    cam = np.array([float(f), float(f) * rng.uniform(0.9, 1.1), w / 2 + 0.5, h / 2 + 0.5,
                    rng.choice([0.0, -3.7]), rng.choice([0.0, 1.3])])
    dt = float(rng.choice([1 / 30, 0.1, 1 / 15, 0.25]))
    dmin, dmax = (0.0, 128.0) if not exotic else (float(rng.choice([0.0, -2.0])), float(rng.choice([128.0, 64.0])))
    return disp, prev, flow, cam, tr, dt, float(f), float(T), dmin, dmax


@pytest.mark.parametrize("seed", range(12))
def test_scene_flow_oracle_equals_numpy_restatement(oracle, seed):
    rng = np.random.RandomState(100 + seed)
    w, h = int(rng.randint(5, 70)), int(rng.randint(5, 50))
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, w, h, exotic=seed % 2 == 1)
    thr = int(rng.choice([0, 1, 5, 100]))
    kw = dict(dynamic_flow_diff=thr, f=f, T=T, min_disparity=dmin, max_disparity=dmax)
    a = oracle.scene_flow(disp, prev, flow, cam, tr, dt, want_depth=True, want_static=True, want_class=True, **kw)
    b = npr.scene_flow(disp, prev, flow, cam, tr, dt, **kw)
    assert_cloud_bit_exact(a["points32"], b["points32"], "cloud")
    assert_cloud_bit_exact(a["depth"][..., None], b["depth"][..., None], "depth")
    assert_cloud_bit_exact(a["static_flow"], b["static_flow"], "static flow")
    assert np.array_equal(a["class"], b["class"])


def test_scene_flow_row_stride_is_honoured(oracle):
    """image.step larger than the row (cv::Mat_ built with msg.image.step, disparity_image_processor.cpp:5)."""
    rng = np.random.RandomState(5)
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, 33, 21, exotic=False)
    kw = dict(f=f, T=T, min_disparity=dmin, max_disparity=dmax)
    dense = oracle.scene_flow(disp, prev, flow, cam, tr, dt, **kw)["points32"]
    pad_d = np.full((21, 40), np.float32(7.0))
    pad_p = np.full((21, 48), np.float32(9.0))
    pad_f = np.full((21, 37, 2), np.float32(3.0))
    pad_d[:, :33], pad_p[:, :33], pad_f[:, :33] = disp, prev, flow
    strided = oracle.scene_flow(pad_d[:, :33], pad_p[:, :33], pad_f[:, :33], cam, tr, dt, **kw)["points32"]
    assert_cloud_bit_exact(strided, dense)


def test_static_scene_identity_ego_has_no_movers(oracle):
    """Property: same disparity in both frames, zero flow, identity ego-motion -> static flow 0, no class IV pixel,
    every chained pixel has exactly zero velocity."""
    rng = np.random.RandomState(6)
    h, w = 40, 56
    disp = rng.uniform(1.0, 100.0, (h, w)).astype(np.float32)
    flow = np.zeros((h, w, 2), np.float32)
    cam = (80.0, 80.0, 28.5, 20.5, 0.0, 0.0)
    r = oracle.scene_flow(disp, disp.copy(), flow, cam, (0, 0, 0, 0, 0, 0, 1), 0.1, f=80.0, T=0.1, want_class=True,
                          want_static=True)
    assert (r["class"] == 2).all()
    assert (r["points32"][..., 4:7] == 0).all()
    assert np.abs(r["static_flow"]).max() < 1e-3


def test_gather_bounds_are_the_only_guard(oracle):
    """A flow that points outside the previous image ends the chain (class II), never reads out of bounds."""
    h, w = 8, 8
    disp = np.full((h, w), 10.0, np.float32)
    for fx, fy in ((100.0, 0.0), (-100.0, 0.0), (0.0, 100.0), (0.0, -100.0), (np.inf, 0.0), (3e9, 0.0)):
        flow = np.zeros((h, w, 2), np.float32)
        flow[..., 0], flow[..., 1] = fx, fy
        r = oracle.scene_flow(disp, disp.copy(), flow, (50.0, 50.0, 4.5, 4.5, 0, 0), (0, 0, 0, 0, 0, 0, 1), 0.1, f=50.0,
                              T=0.1, want_class=True)
        assert (r["class"] == 1).all(), (fx, fy)


# ---- clusterer -----------------------------------------------------------------------------------------------------
def _blob_cloud(rng, w, h, n_blobs, with_depth_steps):
    pts = np.zeros((h, w, 8), np.float32)
    pts[..., 2] = 5.0
    for _ in range(n_blobs):
        bw, bh = rng.randint(2, 14), rng.randint(2, 12)
        x0, y0 = rng.randint(0, w - bw), rng.randint(0, h - bh)
        sp = rng.uniform(0.31, 2.0)
        pts[y0:y0 + bh, x0:x0 + bw, 4] = sp * rng.choice([-1, 1])
        pts[y0:y0 + bh, x0:x0 + bw, 5] = rng.uniform(-0.2, 0.2, (bh, bw))
        if with_depth_steps:
            pts[y0:y0 + bh, x0:x0 + bw, 2] += rng.choice([0.0, 0.1, 0.2, 1.0])
    speck = rng.rand(h, w) < 0.03
    pts[speck, 6] = 1.0
    hole = rng.rand(h, w) < 0.02
    pts[hole] = np.nan
    pts[..., 0] = pts[..., 2] * (np.arange(w, dtype=np.float32)[None, :] - w / 2) / 60
    pts[..., 1] = pts[..., 2] * (np.arange(h, dtype=np.float32)[:, None] - h / 2) / 60
    return pts


@pytest.mark.parametrize("seed", range(10))
def test_cluster_oracle_equals_python_restatement(oracle, seed):
    rng = np.random.RandomState(200 + seed)
    w, h = int(rng.randint(20, 64)), int(rng.randint(16, 48))
    pts = _blob_cloud(rng, w, h, n_blobs=int(rng.randint(2, 9)), with_depth_steps=seed % 2 == 0)
    kw = dict(cluster_size=int(rng.choice([2, 10, 40])), depth_diff=0.15, dynamic_speed=0.3,
              neighbor_distance=int(rng.choice([1, 2, 4, 10])))
    a = oracle.cluster(pts, **kw)
    b = npr.cluster(pts, **kw)
    assert np.array_equal(a["labels"], b["labels"])
    assert a["n_objects"] == len(b["objects"])
    for x, y in zip(a["objects"], b["objects"]):
        assert x["id"] == y["id"] and x["n_points"] == y["n_points"]
        assert x["bbox"].tobytes() == y["bbox"].tobytes() and x["center"].tobytes() == y["center"].tobytes()


def test_cluster_anti_diagonal_contact_does_not_connect_at_n1(oracle):
    """The window is du,dv in [-N,0]: a pixel's up-RIGHT neighbour is never probed, so two pixels touching only on
    the anti-diagonal are not linked at N=1 (they are at N>=2 through a shared up-left window only if a third
    dynamic pixel bridges them)."""
    pts = np.zeros((6, 6, 8), np.float32)
    pts[..., 2] = 3.0
    pts[2, 3, 4] = 1.0   # (u=3,v=2)
    pts[3, 2, 4] = 1.0   # (u=2,v=3): up-right neighbour of it is (3,2)
    r = oracle.cluster(pts, cluster_size=2, neighbor_distance=1)
    assert (r["labels"] == -1).all()
    pts[2, 2, 4] = 1.0   # bridge: (2,2) is left of (3,2) and above (2,3)
    r = oracle.cluster(pts, cluster_size=2, neighbor_distance=1)
    assert (r["labels"] >= 0).sum() == 3 and r["n_objects"] == 1


def test_cluster_isolated_dynamic_pixels_stay_unlabelled(oracle):
    pts = np.zeros((10, 10, 8), np.float32)
    pts[..., 2] = 3.0
    pts[::5, ::5, 4] = 1.0
    r = oracle.cluster(pts, cluster_size=1, neighbor_distance=4)
    assert (r["labels"] == -1).all() and r["n_objects"] == 0 and r["dynamic"].sum() == 4


def test_cluster_depth_gate_splits(oracle):
    pts = np.zeros((8, 16, 8), np.float32)
    pts[..., 4] = 1.0
    pts[:, :8, 2] = 3.0
    pts[:, 8:, 2] = 3.5      # 0.5 m step > depth_diff
    r = oracle.cluster(pts, cluster_size=10, neighbor_distance=4)
    assert r["n_objects"] == 2
    assert len(np.unique(r["labels"][:, :8])) == 1 and len(np.unique(r["labels"][:, 8:])) == 1
    assert r["labels"][0, 0] == 0 and r["labels"][0, 8] == 1   # ids in raster order of first accepted pair


def test_cluster_size_threshold_is_inclusive(oracle):
    pts = np.zeros((8, 30, 8), np.float32)
    pts[..., 2] = 3.0
    pts[1:4, 1:5, 4] = 1.0      # 12 pixels
    pts[1:4, 10:14, 4] = 1.0    # 12 pixels
    pts[5, 10, 4] = 0.0
    pts[1:4, 20:23, 4] = 1.0    # 9 pixels
    r = oracle.cluster(pts, cluster_size=12, neighbor_distance=1)
    assert r["n_objects"] == 2 and [o["n_points"] for o in r["objects"]] == [12, 12]
    r = oracle.cluster(pts, cluster_size=13, neighbor_distance=1)
    assert r["n_objects"] == 0


def test_clusters_image_and_member_order(oracle):
    """publishClustersImage / ColorSet (clusterer_nodelet.cpp:293-322, color_set.cpp:16-26) and the u-outer / v-inner
    member order of clusterMap2IndicesCluster (:97-117) on a hand-made label map."""
    import os
    lut = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hsv_colormap_bgr.npy"))
    labels = np.full((4, 5), -1, np.int32)
    labels[0, 1:3] = 0
    labels[1:4, 3] = 1
    labels[3, 0] = 2
    labels[1, 1] = 0
    img = oracle.clusters_image(labels, 3, lut)
    assert (img[labels == -1] == 0).all()
    for k in range(3):
        assert (img[labels == k] == lut[k * 255 // 3]).all()
    members = oracle.cluster_members(labels, 3)
    assert members[0].tolist() == [1, 6, 2]          # column 1 (rows 0, 1) before column 2
    assert members[1].tolist() == [8, 13, 18] and members[2].tolist() == [15]
    assert (oracle.clusters_image(labels * 0 - 1, 0, lut) == 0).all()   # no clusters: ColorSet::resize returns early
    with pytest.raises(RuntimeError):
        oracle.clusters_image(labels, 2, lut)        # label 2 with a 2-entry colour set: cv::Mat::at throws


def test_embedded_colormap_equals_the_cv2_fixture(mod):
    """The table compiled into libmod_b200.so is the fixture generated from cv2 (and equals cv2's if it is importable)."""
    import os
    from moving_object_detector_b200.api import hsv_colormap
    lut = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hsv_colormap_bgr.npy"))
    assert np.array_equal(hsv_colormap(), lut)
    try:
        import cv2
    except ImportError:
        return
    live = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_HSV)[0]
    assert np.array_equal(live, lut)
