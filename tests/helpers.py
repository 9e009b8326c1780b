"""Shared helpers of the parity tests: bit-exact comparison of clouds, partitions and object lists."""
import numpy as np

from moving_object_detector_b200 import CameraInfo, DisparityImage, TransformMsg


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_cloud_bit_exact(got, want, what="cloud"):
    """Both (H,W,8) float32.  NaN payloads are normalised (any NaN equals any NaN), everything else bit-exact."""
    got = np.ascontiguousarray(got, np.float32)
    want = np.ascontiguousarray(want, np.float32)
    assert got.shape == want.shape
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), f"{what}: NaN pattern differs at {np.argwhere(gn != wn)[:5]}"
    same = (bits(got) == bits(want)) | gn
    if not same.all():
        idx = np.argwhere(~same)
        y, x, c = idx[0]
        raise AssertionError(f"{what}: {len(idx)} values differ, first at (y={y},x={x},field={c}): "
                             f"got {got[y, x, c]!r} want {want[y, x, c]!r}")


def seq_inputs(seq, i):
    """Message-like inputs of pair i of a synth sequence (prev = frame i, now = frame i+1)."""
    disp = seq["disp"].cpu().numpy() if hasattr(seq["disp"], "cpu") else seq["disp"]
    flows = seq["flows"].cpu().numpy() if hasattr(seq["flows"], "cpu") else seq["flows"]
    mk = lambda img: DisparityImage(np.ascontiguousarray(img), seq["f"], seq["T"], seq["min_disparity"],
                                    seq["max_disparity"])
    cam = seq["cam"]
    t = seq["transforms"][i]
    return dict(now=mk(disp[i + 1]), prev=mk(disp[i]), flow=np.ascontiguousarray(flows[i]),
                cam=CameraInfo.pinhole(cam[0], cam[1], cam[2], cam[3], cam[4], cam[5]),
                prev2now=TransformMsg(t[:3], t[3:]), dt_sec=seq["dts"][i])


def oracle_scene_flow(oracle, seq, i, dynamic_flow_diff=5, **kw):
    disp = seq["disp"].cpu().numpy() if hasattr(seq["disp"], "cpu") else seq["disp"]
    flows = seq["flows"].cpu().numpy() if hasattr(seq["flows"], "cpu") else seq["flows"]
    return oracle.scene_flow(np.ascontiguousarray(disp[i + 1]), np.ascontiguousarray(disp[i]),
                             np.ascontiguousarray(flows[i]), seq["cam"], seq["transforms"][i], seq["dts"][i],
                             dynamic_flow_diff=dynamic_flow_diff, f=seq["f"], T=seq["T"],
                             min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"], **kw)


def assert_same_partition(a, b):
    """Label images equal up to a permutation of the non-negative labels."""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    assert np.array_equal(a < 0, b < 0), "unlabelled sets differ"
    m = a >= 0
    pairs = np.unique(np.stack([a[m], b[m]]), axis=1)
    assert len(np.unique(pairs[0])) == pairs.shape[1] and len(np.unique(pairs[1])) == pairs.shape[1], \
        "label images are not a bijective relabelling of each other"


def assert_objects_match(got, want, points32, labels, tol=0.0):
    """Objects in id order: id, n_points, bbox, centre bit-exact; velocity's norm equals the oracle's and the vector is
    one of the cluster members tied at that norm (std::sort leaves ties unspecified, clusterer_nodelet.cpp:168-174)."""
    assert len(got) == len(want), f"{len(got)} objects vs {len(want)}"
    pts = np.asarray(points32).reshape(-1, 8)
    lab = np.asarray(labels).ravel()
    for g, w in zip(got, want):
        assert g["id"] == w["id"]
        assert g["n_points"] == w["n_points"]
        assert np.array_equal(bits(g["bbox"]), bits(w["bbox"])), (g["bbox"], w["bbox"])
        assert np.array_equal(bits(g["center"]), bits(w["center"])), (g["center"], w["center"])
        vg, vw = g["velocity"].astype(np.float32), w["velocity"].astype(np.float32)
        ng = np.sqrt((vg[0] * vg[0] + vg[1] * vg[1]) + vg[2] * vg[2])
        nw = np.sqrt((vw[0] * vw[0] + vw[1] * vw[1]) + vw[2] * vw[2])
        assert ng == nw, f"median speed differs: {ng} vs {nw}"
        members = pts[lab == g["id"]][:, 4:7]
        assert (members == vg[None, :]).all(axis=1).any(), "velocity is not a member of the cluster"


def foreign_cloud_with_nonfinite_depth(w=320, h=200, seed=77, cluster_size=100):
    """A clusterer-stress cloud as another publisher might send it: some DYNAMIC points carry z = NaN / +inf / -inf
    (x, y alike for a few), on top of depth noise of the order of depth_diff.  comparePoints' gate is
    "if (depthDiff > th) return" (clusterer_nodelet.cpp:205-209), so a NaN difference PASSES and an infinite one fails
    unless the threshold is infinite too -- the tile kernels stage non-dynamic cells as z = +inf and must take their
    explicit-flag path for such points and thresholds."""
    from moving_object_detector_b200 import synth
    cloud, _ = synth.make_stress_cloud(w, h, seed=seed, cluster_size=cluster_size, fx=140.0)
    rng = np.random.RandomState(seed)
    cloud[..., 2] += rng.normal(0.0, 0.08, (h, w)).astype(np.float32)
    v = cloud[..., 4:7]
    dyn = np.sqrt((v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]) >= 0.3
    pick = rng.rand(h, w)
    cloud[dyn & (pick < 0.03), 2] = np.nan
    cloud[dyn & (pick >= 0.03) & (pick < 0.04), 2] = np.inf
    cloud[dyn & (pick >= 0.04) & (pick < 0.05), 2] = -np.inf
    cloud[dyn & (pick >= 0.05) & (pick < 0.055), 0] = np.nan
    return cloud
