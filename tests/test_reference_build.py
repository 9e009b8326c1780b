"""Pins the CPU oracle to the REFERENCE'S OWN CODE.

oracle/_ref/libmod_ref.so holds the reference's unmodified translation units (disparity_image_processor.cpp,
scene_flow_constructor.cpp, clusterer_nodelet.cpp, color_set.cpp, lookup_table.cpp) compiled from /root/reference
against stand-in dependency headers (oracle/ref_shim/ref_shim.h; oracle/Makefile target ref_full).  These tests run
SceneFlowConstructor::construct() and ClustererNodelet::dataCB() themselves and require the oracle restatement
(oracle/mod_oracle.cpp) to equal what the two nodes publish, bit for bit: cloud, ~depth, ~synthetic_optical_flow,
dynamic map, cluster map with the reference's ids, MovingObjectArray, ~clusters_image, ~clusters markers, and the
degrade rules / subscriber gating around them.

CPU only; skipped when the library was not built (it needs /root/reference at build time)."""
import dataclasses
import glob
import os

import numpy as np
import pytest

from helpers import assert_cloud_bit_exact
from test_oracle import _blob_cloud, _random_pair

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = [0, 1, 2, 4, 5, 6]   # x y z vx vy vz; the two pad floats of the 32-byte record are uninitialised upstream


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not built (needs /root/reference at build time)")
    ref_py.load()
    return ref_py


def _compare_scene_flow(oracle, ref, disp, prev, flow, cam, tr, dt_wish, **kw):
    sp, sn, dt = ref.stamps_for(dt_wish)
    o = oracle.scene_flow(disp, prev, flow, cam, tr, dt, want_depth=True, want_static=True, want_class=True, **kw)
    r = ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, **kw)
    assert r["status"] == 0 and r["published"] == 7
    assert_cloud_bit_exact(o["points32"][..., FIELDS], r["points32"][..., FIELDS], "cloud")
    assert_cloud_bit_exact(o["depth"][..., None], r["depth"][..., None], "depth")
    assert_cloud_bit_exact(o["static_flow"], r["static_flow"], "static flow")
    return o, r


def _compare_cluster(oracle, ref, pts, **kw):
    o = oracle.cluster(pts, **kw)
    r = ref.cluster(pts, **kw)
    assert r["status"] == 0 and r["published"] == 15 and r["objects_lossless"]
    assert np.array_equal(o["dynamic"], r["dynamic"]), "dynamic map"
    assert np.array_equal(o["labels"], r["labels"]), "cluster map (ids in the reference's order)"
    assert o["n_objects"] == r["n_objects"]
    for k in ("id", "center", "bbox"):
        assert o["objects"][k].tobytes() == r["objects"][k].tobytes(), k
    # orientation is the identity quaternion (clusterer_nodelet.cpp:163-166)
    assert (r["objects_raw"][:, 4:8] == [0, 0, 0, 1]).all()
    # velocity: the member the reference's std::sort leaves at n/2; ties may be ordered differently by another sort,
    # so the norm must agree exactly and the vector must belong to the cluster
    vo, vr = o["objects"]["velocity"], r["objects"]["velocity"]
    no = np.sqrt((vo[:, 0] * vo[:, 0] + vo[:, 1] * vo[:, 1]) + vo[:, 2] * vo[:, 2])
    nr = np.sqrt((vr[:, 0] * vr[:, 0] + vr[:, 1] * vr[:, 1]) + vr[:, 2] * vr[:, 2])
    assert np.array_equal(no, nr)
    for k, obj in enumerate(o["objects"]):
        assert obj["n_points"] == int((r["labels"] == k).sum())
    nc = r["number_of_clusters"]
    assert nc == (o["labels"].max() + 1 if (o["labels"] >= 0).any() else 0)
    lut = np.load(os.path.join(GOLDEN, "hsv_colormap_bgr.npy"))
    if nc > 0:
        assert np.array_equal(oracle.clusters_image(o["labels"], nc, lut), r["clusters_image"]), "~clusters_image"
        members = oracle.cluster_members(o["labels"], nc)
        flat = np.ascontiguousarray(pts, np.float32).reshape(-1, 8)
        assert r["n_markers"] == nc
        for k, (idx, got) in enumerate(zip(members, r["markers"])):
            assert np.array_equal(flat[idx][:, :3].astype(np.float64), got), f"marker {k} point order"
            b, g, rr = lut[k * 255 // nc]
            assert np.array_equal(r["marker_rgba"][k], np.array([rr / 255.0, g / 255.0, b / 255.0, 1.0], np.float32))
    else:
        assert (r["clusters_image"] == 0).all() and r["n_markers"] == 0
    return o, r


def test_reconfigure_defaults_are_the_cfg_defaults(ref):
    """cfg/SceneFlowConstructor.cfg:8, cfg/Clusterer.cfg:8-13 as latched by the nodes' reconfigureCB at start-up."""
    assert ref.defaults() == dict(dynamic_flow_diff=5, cluster_size=2500, neighbor_distance=4, depth_diff=0.15,
                                  dynamic_speed=0.3)
    from moving_object_detector_b200 import ClustererConfig, SceneFlowConstructorConfig
    c, s = ClustererConfig(), SceneFlowConstructorConfig()
    assert (s.dynamic_flow_diff, c.cluster_size, c.neighbor_distance, c.depth_diff, c.dynamic_speed) == \
        (5, 2500, 4, 0.15, 0.3)


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "sf_*.npz"))))
def test_reference_reproduces_scene_flow_golden(oracle, ref, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    f, T, dmin, dmax = (float(v) for v in g["meta"])
    sp, sn, dt = ref.stamps_for(float(g["dt"]))
    assert dt == float(g["dt"]), "fixture dt must be expressible as a difference of ROS stamps"
    r = ref.scene_flow(g["disp_now"], g["disp_prev"], g["flow"], g["cam"], g["transform"], sp, sn,
                       dynamic_flow_diff=int(g["dynamic_flow_diff"]), f=f, T=T, min_disparity=dmin, max_disparity=dmax)
    assert r["status"] == 0 and r["published"] == 7
    assert_cloud_bit_exact(r["points32"][..., FIELDS], g["points32"][..., FIELDS], "cloud")
    assert_cloud_bit_exact(r["depth"][..., None], g["depth"][..., None], "depth")
    assert_cloud_bit_exact(r["static_flow"], g["static_flow"], "static flow")


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "cl_*.npz"))))
def test_reference_reproduces_cluster_golden(oracle, ref, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    kw = dict(cluster_size=int(g["params"][0]), neighbor_distance=int(g["params"][1]),
              depth_diff=float(g["thresholds"][0]), dynamic_speed=float(g["thresholds"][1]))
    o, r = _compare_cluster(oracle, ref, g["points32"], **kw)
    assert np.array_equal(r["labels"], g["labels"]) and np.array_equal(r["dynamic"], g["dynamic"])
    for k in ("id", "center", "bbox", "velocity"):
        assert r["objects"][k].tobytes() == g["objects"][k].tobytes(), k


@pytest.mark.parametrize("seed", range(16))
def test_scene_flow_oracle_equals_reference_random(oracle, ref, seed):
    """Randomised pairs; odd seeds carry every invalid encoding (0, -0, <0, NaN, +-inf, >max, subnormal), flows
    that round at .5 / overflow int, un-normalised quaternions, Tx,Ty != 0, non-default disparity ranges."""
    rng = np.random.RandomState(100 + seed)
    w, h = int(rng.randint(5, 70)), int(rng.randint(5, 50))
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, w, h, exotic=seed % 2 == 1)
    thr = int(rng.choice([0, 1, 5, 100]))
    _compare_scene_flow(oracle, ref, disp, prev, flow, cam, tr, dt, dynamic_flow_diff=thr, f=f, T=T, min_disparity=dmin,
                        max_disparity=dmax)


@pytest.mark.parametrize("dt", [0.0, 1e-9, 2.5, 1 / 15])
def test_scene_flow_stamp_differences(oracle, ref, dt):
    """dt comes from (stamp_now - stamp_prev).toSec(); equal stamps divide by zero (inf / NaN velocities)."""
    rng = np.random.RandomState(77)
    disp, prev, flow, cam, tr, _, f, T, dmin, dmax = _random_pair(rng, 40, 30, exotic=False)
    _compare_scene_flow(oracle, ref, disp, prev, flow, cam, tr, dt, dynamic_flow_diff=1, f=f, T=T, min_disparity=dmin,
                        max_disparity=dmax)


def test_scene_flow_row_strides(oracle, ref):
    """disparity image.step and the flow cv::Mat step larger than a row (disparity_image_processor.cpp:5)."""
    rng = np.random.RandomState(5)
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, 33, 21, exotic=False)
    pad_d = np.full((21, 40), np.float32(7.0))
    pad_p = np.full((21, 48), np.float32(9.0))
    pad_f = np.full((21, 37, 2), np.float32(3.0))
    pad_d[:, :33], pad_p[:, :33], pad_f[:, :33] = disp, prev, flow
    kw = dict(f=f, T=T, min_disparity=dmin, max_disparity=dmax)
    o, r = _compare_scene_flow(oracle, ref, pad_d[:, :33], pad_p[:, :33], pad_f[:, :33], cam, tr, dt, **kw)
    dense = ref.scene_flow(disp, prev, flow, cam, tr, *ref.stamps_for(dt)[:2], **kw)
    assert_cloud_bit_exact(dense["points32"][..., FIELDS], r["points32"][..., FIELDS])


def test_point_cloud2_wire_format(ref):
    """pcl::toROSMsg of PointCloud<PointXYZVelocity> (scene_flow_constructor.cpp:351-362, pcl_point_xyz_velocity.h)."""
    rng = np.random.RandomState(3)
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, 24, 10, exotic=False)
    sp, sn, _ = ref.stamps_for(dt)
    m = ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, f=f, T=T)["cloud_meta"]
    width, height, point_step, row_step, is_dense, big_endian, n_fields, names_ok = m[:8]
    assert (width, height, point_step, row_step, big_endian, n_fields, names_ok) == (24, 10, 32, 32 * 24, 0, 6, 1)
    assert is_dense == 1   # PointCloud(w, h, value) leaves is_dense = true although the cloud holds NaNs
    assert m[8:20].reshape(6, 2).tolist() == [[0, 7], [4, 7], [8, 7], [16, 7], [20, 7], [24, 7]]   # FLOAT32 = 7
    from moving_object_detector_b200 import api
    if hasattr(api, "POINT_FIELDS"):
        assert [(n, o) for n, o, *_ in api.POINT_FIELDS] == [("x", 0), ("y", 4), ("z", 8), ("vx", 16), ("vy", 20),
                                                             ("vz", 24)]


def test_degrade_rules_and_subscriber_gating(ref):
    """construct() (scene_flow_constructor.cpp:91-147): ~optical_flow is echoed first; ~depth needs only the current
    disparity; without a flow nothing else happens; without the previous disparity or the ego-motion transform no
    cloud is published; every output is gated on its subscriber count."""
    rng = np.random.RandomState(9)
    disp, prev, flow, cam, tr, dt, f, T, dmin, dmax = _random_pair(rng, 20, 12, exotic=False)
    sp, sn, _ = ref.stamps_for(dt)
    kw = dict(f=f, T=T)
    assert ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, subscribers=15, **kw)["published"] == 15
    assert ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, subscribers=0, **kw)["published"] == 0
    assert ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, subscribers=1, **kw)["published"] == 1
    r = ref.scene_flow(disp, prev, None, cam, tr, sp, sn, subscribers=15, **kw)
    assert r["status"] == 0 and r["published"] == 2                       # depth only
    assert ref.scene_flow(disp, None, flow, cam, tr, sp, sn, subscribers=15, **kw)["published"] == 2 | 8
    assert ref.scene_flow(disp, prev, flow, cam, None, sp, sn, subscribers=15, **kw)["published"] == 2 | 8
    assert ref.scene_flow(None, prev, flow, cam, tr, sp, sn, subscribers=15, **kw)["published"] == 8
    echo = ref.scene_flow(disp, prev, flow, cam, tr, sp, sn, subscribers=8, **kw)
    assert echo["published"] == 8 and np.array_equal(echo["flow_echo"], flow)
    none = ref.cluster(np.zeros((4, 4, 8), np.float32), subscribers=0)
    assert none["status"] == 0 and none["published"] == 0                  # clustering still runs, nothing is sent


@pytest.mark.parametrize("seed", range(12))
def test_cluster_oracle_equals_reference_random(oracle, ref, seed):
    rng = np.random.RandomState(200 + seed)
    w, h = int(rng.randint(20, 64)), int(rng.randint(16, 48))
    pts = _blob_cloud(rng, w, h, n_blobs=int(rng.randint(2, 9)), with_depth_steps=seed % 2 == 0)
    _compare_cluster(oracle, ref, pts, cluster_size=int(rng.choice([2, 10, 40])), depth_diff=0.15, dynamic_speed=0.3,
                     neighbor_distance=int(rng.choice([1, 2, 4, 10])))


def test_cluster_speed_ties_pick_the_same_member(oracle, ref):
    """Many members share |v| exactly (std::sort is unstable): norm and membership are what is defined."""
    pts = np.zeros((24, 40, 8), np.float32)
    pts[..., 2] = 4.0
    pts[2:20, 3:30, 4] = 0.5
    pts[2:20, 3:30, 5] = np.where(np.arange(27)[None, :] % 2 == 0, 0.25, -0.25)
    pts[5:9, 5:9, 4] = 0.9
    _compare_cluster(oracle, ref, pts, cluster_size=10, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=2)


def test_cluster_median_below_speed_threshold_is_rejected_and_ids_stay_dense(oracle, ref):
    """cluster2MovingObject returns false when |median v| < dynamic_speed (never true for the reference's own clouds,
    reachable through NaN speeds); ids count accepted objects only (clusterer_nodelet.cpp:324-343)."""
    pts = np.zeros((16, 48, 8), np.float32)
    pts[..., 2] = 4.0
    pts[2:12, 2:12, 4] = 1.0
    pts[2:12, 20:30, 4] = 0.3       # float(0.3) >= double 0.3: dynamic (0.3f widens to 0.30000001192...)
    pts[2:12, 36:46, 4] = 0.30000004
    o, r = _compare_cluster(oracle, ref, pts, cluster_size=50, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=1)
    assert r["n_objects"] == 3 and r["objects"]["id"].tolist() == [0, 1, 2]


def test_pipeline_through_the_published_message(oracle, ref):
    """640x480 Gazebo-like pair (BASELINE.json configs[0]): construct() publishes ~scene_flow, the clusterer nodelet
    receives that very PointCloud2 (pcl::toROSMsg -> fromROSMsg) and publishes ~moving_objects."""
    from moving_object_detector_b200 import synth
    seq = synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)
    disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
    kw = dict(f=seq["f"], T=seq["T"], min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"])
    o, r = _compare_scene_flow(oracle, ref, np.ascontiguousarray(disp[1]), np.ascontiguousarray(disp[0]),
                               np.ascontiguousarray(flows[0]), seq["cam"], seq["transforms"][0], seq["dts"][0], **kw)
    counts = np.bincount(o["class"].ravel(), minlength=4)
    assert (counts > 0).all()
    rc = ref.cluster(None, shape=disp[0].shape)
    oc = oracle.cluster(o["points32"])
    assert rc["status"] == 0 and rc["n_objects"] == oc["n_objects"] >= 2
    assert np.array_equal(rc["labels"], oc["labels"]) and np.array_equal(rc["dynamic"], oc["dynamic"])
    for k in ("id", "center", "bbox", "velocity"):
        assert rc["objects"][k].tobytes() == oc["objects"][k].tobytes(), k


@pytest.mark.parametrize("depth_diff", [0.15, float("inf"), 1e30, 0.0])
def test_nonfinite_depths_oracle_equals_reference_labels(oracle, ref, depth_diff):
    """Dynamic points with z = NaN / +-inf and thresholds up to +inf: dynamic map and cluster map (the reference's ids).
    Objects are not compared: pcl::getMinMax3D on NaN coordinates depends on the member order."""
    from helpers import foreign_cloud_with_nonfinite_depth
    cloud = foreign_cloud_with_nonfinite_depth()
    for n in (1, 4, 7, 10):
        kw = dict(cluster_size=100, depth_diff=depth_diff, dynamic_speed=0.3, neighbor_distance=n)
        o, r = oracle.cluster(cloud, **kw), ref.cluster(cloud, **kw)
        assert r["status"] == 0
        assert np.array_equal(o["dynamic"], r["dynamic"]), ("dynamic map", n)
        assert np.array_equal(o["labels"], r["labels"]), ("cluster map", n)
        assert o["n_objects"] == r["n_objects"] and (depth_diff == 0.0 or o["n_objects"] > 0)


def test_stress_cloud_oracle_equals_reference(oracle, ref):
    """Clusterer-stress structure (serpentine, blobs straddling cluster_size, specks, checkerboards, depth steps)."""
    from moving_object_detector_b200 import synth
    cloud, _ = synth.make_stress_cloud(320, 200, seed=4104, cluster_size=100, fx=140.0)
    for n in (1, 3, 4, 7, 10):
        _compare_cluster(oracle, ref, cloud, cluster_size=100, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=n)
