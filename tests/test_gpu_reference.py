"""GPU parity against the REFERENCE'S OWN CODE: the CUDA path (through the C ABI) vs oracle/_ref/libmod_ref.so, the
reference's unmodified translation units compiled against stand-in dependency headers (oracle/Makefile ref_full).
The prebuilt library travels to the GPU box with the snapshot; /root/reference itself is not read here.

What is compared is what the two ROS nodes publish: ~scene_flow (PointCloud2 payload), ~depth,
~synthetic_optical_flow, ~moving_objects, the cluster map, ~clusters_image, ~clusters marker point order."""
import dataclasses
import os

import numpy as np
import pytest

from helpers import assert_cloud_bit_exact, assert_objects_match, seq_inputs
from moving_object_detector_b200 import ClustererConfig, Context, SceneFlowConstructorConfig, synth

pytestmark = pytest.mark.gpu
FIELDS = [0, 1, 2, 4, 5, 6]


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not in the snapshot (built only where /root/reference exists)")
    ref_py.load()
    return ref_py


@pytest.fixture(scope="module")
def ctx():
    c = Context(2208, 1242, max_batch=2)
    yield c
    c.close()


def _ref_pair(ref, seq, i, thr=5, subscribers=7):
    disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
    sp, sn, dt = ref.stamps_for(seq["dts"][i])
    r = ref.scene_flow(np.ascontiguousarray(disp[i + 1]), np.ascontiguousarray(disp[i]), np.ascontiguousarray(flows[i]),
                       seq["cam"], seq["transforms"][i], sp, sn, dynamic_flow_diff=thr, f=seq["f"], T=seq["T"],
                       min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"], subscribers=subscribers)
    assert r["status"] == 0
    return r, dt


def _ref_objects(r):
    objs = r["objects"].copy()
    objs["n_points"] = [int((r["labels"] == k).sum()) for k in objs["id"]]
    return objs


@pytest.mark.parametrize("name,pairs", [("gazebo_640x480", 2), ("zed_1280x720", 2), ("kitti_1242x375", 1)])
def test_constructor_and_clusterer_equal_the_reference_nodes(ctx, ref, name, pairs):
    seq = synth.make_sequence(synth.CONFIGS[name], pairs + 1)
    for i in range(pairs):
        r, dt = _ref_pair(ref, seq, i)
        inp = dict(seq_inputs(seq, i), dt_sec=dt)
        got, depth, static = ctx.scene_flow(**inp, want_depth=True, want_static_flow=True)
        assert_cloud_bit_exact(got[..., FIELDS], r["points32"][..., FIELDS], "~scene_flow")
        assert_cloud_bit_exact(depth[..., None], r["depth"][..., None], "~depth")
        assert_cloud_bit_exact(static, r["static_flow"], "~synthetic_optical_flow")
        got_f, _, _ = ctx.scene_flow(**inp)      # the filtered-gate kernel (no static-flow image requested)
        assert_cloud_bit_exact(got_f[..., FIELDS], r["points32"][..., FIELDS], "~scene_flow (filtered gate)")

        rc = ref.cluster(None, shape=got.shape[:2])   # the nodelet consumes the message the constructor published
        assert rc["status"] == 0 and rc["objects_lossless"]
        res = ctx.pipeline(**inp, want_points=True, want_labels=True)
        assert_cloud_bit_exact(res.points32[..., FIELDS], r["points32"][..., FIELDS])
        assert np.array_equal(res.labels, rc["labels"]), "cluster map (ids in the reference's order)"
        assert rc["n_objects"] >= 1
        assert_objects_match(res.objects, _ref_objects(rc), r["points32"], rc["labels"])
        img = ctx.clusters_image(res.labels, rc["number_of_clusters"])
        assert np.array_equal(img, rc["clusters_image"]), "~clusters_image"


@pytest.mark.parametrize("thr", [1, 100])
def test_dynamic_flow_diff_values(ctx, ref, thr):
    seq = synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)
    r, dt = _ref_pair(ref, seq, 0, thr=thr, subscribers=1)
    got, _, _ = ctx.scene_flow(**dict(seq_inputs(seq, 0), dt_sec=dt),
                               config=SceneFlowConstructorConfig(dynamic_flow_diff=thr))
    assert_cloud_bit_exact(got[..., FIELDS], r["points32"][..., FIELDS])


def test_invalid_encodings_and_edge_flows(ctx, ref):
    """Every invalid disparity encoding on both frames, flows that round at .5, overflow or are NaN/inf."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sf_small_edges.npz"))
    from moving_object_detector_b200 import CameraInfo, DisparityImage, TransformMsg
    f, T, dmin, dmax = (float(v) for v in g["meta"])
    sp, sn, dt = ref.stamps_for(float(g["dt"]))
    r = ref.scene_flow(g["disp_now"], g["disp_prev"], g["flow"], g["cam"], g["transform"], sp, sn, f=f, T=T,
                       min_disparity=dmin, max_disparity=dmax)
    tr = g["transform"]
    got, depth, static = ctx.scene_flow(DisparityImage(g["disp_now"], f, T, dmin, dmax),
                                        DisparityImage(g["disp_prev"], f, T, dmin, dmax), g["flow"],
                                        CameraInfo.pinhole(*g["cam"]), TransformMsg(tr[:3], tr[3:]), dt,
                                        want_depth=True, want_static_flow=True)
    assert_cloud_bit_exact(got[..., FIELDS], r["points32"][..., FIELDS])
    assert_cloud_bit_exact(depth[..., None], r["depth"][..., None])
    assert_cloud_bit_exact(static, r["static_flow"])


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 10])
def test_clusterer_stress_equals_the_reference_nodelet(ctx, ref, n):
    """Clusterer-stress structure (serpentine, blobs straddling cluster_size, specks, checkerboards, depth steps)
    at every kernel family (tile-fast N = 1..4, generic N = 5, 10)."""
    cloud, _ = synth.make_stress_cloud(736, 414, seed=4004, cluster_size=300, fx=466.0)
    rc = ref.cluster(cloud, cluster_size=300, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=n)
    assert rc["status"] == 0 and rc["n_objects"] >= 3
    got = ctx.cluster(cloud, ClustererConfig(300, 0.15, 0.3, n))
    assert np.array_equal(got.labels, rc["labels"])
    assert_objects_match(got.objects, _ref_objects(rc), cloud, rc["labels"])
    assert np.array_equal(ctx.clusters_image(got.labels, rc["number_of_clusters"]), rc["clusters_image"])


def test_hd2k_stress_frame_equals_the_reference_nodelet(ctx, ref):
    """BASELINE.json configs[3]: one 2208x1242 dense-mover frame at the default parameters."""
    cloud, _ = synth.make_stress_cloud(2208, 1242, seed=4004)
    rc = ref.cluster(cloud, subscribers=1)
    assert rc["status"] == 0 and rc["n_objects"] >= 3
    got = ctx.cluster(cloud, ClustererConfig())
    assert np.array_equal(got.labels, rc["labels"])
    assert_objects_match(got.objects, _ref_objects(rc), cloud, rc["labels"])
