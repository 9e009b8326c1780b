"""The drop-in ROS package (moving_object_detector_b200/ros/scene_flow_clusterer): the product's ClustererNodelet,
written against the real ROS API, compiled unmodified against the stand-in ROS headers (oracle/ref_shim) and handed
the same sensor_msgs/PointCloud2 as the REFERENCE's nodelet (oracle/_ref/libmod_ref.so).  What the two publish on
~moving_objects, ~clusters and ~clusters_image must be equal message for message.

CPU: package surface (plugin description, class / topic / parameter names, defaults) and loud failure without a GPU.
GPU (-m gpu): message parity with the reference nodelet."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "moving_object_detector_b200", "ros", "scene_flow_clusterer")


@pytest.fixture(scope="module")
def plugin():
    import ros_pkg_py
    if not ros_pkg_py.available():
        pytest.skip("oracle/_shim/libclusterer_nodelet_b200.so not built (make -C oracle ros_pkg)")
    ros_pkg_py.load()
    return ros_pkg_py


def test_package_surface_matches_the_reference_contract():
    """Plugin name / type / base class, library path, wrapper executable, package name, cfg parameters: the contract of
    SURVEY.md 8(b) (reference: nodelet_plugins.xml:3-8, package.xml:20-22, CMakeLists.txt:52-53, cfg/Clusterer.cfg:8-11)."""
    xml = open(os.path.join(PKG, "nodelet_plugins.xml")).read()
    assert 'path="lib/libscene_flow_clusterer"' in xml
    assert 'name="scene_flow_clusterer/scene_flow_clusterer"' in xml
    assert 'type="scene_flow_clusterer::ClustererNodelet"' in xml and 'base_class_type="nodelet::Nodelet"' in xml
    pkg = open(os.path.join(PKG, "package.xml")).read()
    assert "<name>scene_flow_clusterer</name>" in pkg and 'nodelet plugin="${prefix}/nodelet_plugins.xml"' in pkg
    cm = open(os.path.join(PKG, "CMakeLists.txt")).read()
    assert "add_library(scene_flow_clusterer " in cm and "add_executable(scene_flow_clusterer_node " in cm
    assert "generate_dynamic_reconfigure_options(cfg/Clusterer.cfg)" in cm
    cfg = open(os.path.join(PKG, "cfg", "Clusterer.cfg")).read()
    params = re.findall(r'gen\.add\("(\w+)", (\w+), 0, "[^"]*", ([\d.]+), ([\d.]+), ([\d.]+)\)', cfg)
    assert params == [("cluster_size", "int_t", "2500", "100", "10000"), ("depth_diff", "double_t", "0.15", "0.01", "1.0"),
                      ("dynamic_speed", "double_t", "0.3", "0.01", "1.0"), ("neighbor_distance", "int_t", "4", "1", "10")]
    assert 'gen.generate(PACKAGE, "scene_flow_clusterer", "Clusterer")' in cfg
    src = open(os.path.join(PKG, "src", "clusterer_nodelet.cpp")).read()
    assert "PLUGINLIB_EXPORT_CLASS(scene_flow_clusterer::ClustererNodelet, nodelet::Nodelet)" in src
    for topic in ('"scene_flow", 10', '"moving_objects", 1', '"clusters", 1', '"clusters_image", 1'):
        assert topic in src
    node = open(os.path.join(PKG, "src", "clusterer_node.cpp")).read()
    assert '"scene_flow_clusterer/scene_flow_clusterer"' in node
    # the reference's wrapper (clusterer_node.cpp:8-13): no load / unload services, no remappings forwarded
    assert "loader(false)" in node and "getRemappings" not in node


def test_if_the_reference_is_here_its_surface_is_the_same():
    ref = "/root/reference/scene_flow_clusterer"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    norm = lambda s: re.sub(r"\s+", " ", s)
    ours, theirs = (norm(open(os.path.join(d, "nodelet_plugins.xml")).read()) for d in (PKG, ref))
    for attr in ("path", "name", "type", "base_class_type"):
        assert re.search(attr + r'="([^"]+)"', ours).group(1) == re.search(attr + r'="([^"]+)"', theirs).group(1)
    grab = lambda s: re.findall(r'gen\.add\("(\w+)", (\w+), 0, "[^"]*", ([\d.]+), ([\d.]+), ([\d.]+)\)', s)
    assert grab(open(os.path.join(PKG, "cfg", "Clusterer.cfg")).read()) == \
        grab(open(os.path.join(ref, "cfg", "Clusterer.cfg")).read())


def test_nodelet_latches_cfg_defaults_and_advertises_the_topics(plugin):
    assert plugin.defaults() == dict(cluster_size=2500, neighbor_distance=4, topics_ok=3, depth_diff=0.15,
                                     dynamic_speed=0.3)


def test_nodelet_drops_the_cloud_loudly_without_a_gpu(plugin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = plugin.cluster(np.zeros((8, 8, 8), np.float32))
    assert r["status"] == 0 and r["published"] == 0   # no CPU fallback: nothing is published


# ---- GPU: message parity with the reference nodelet ----------------------------------------------------------------
def _compare(plugin, ref, cloud, **kw):
    a = plugin.cluster(cloud, **kw)
    b = ref.cluster(cloud, **kw)
    assert a["status"] == 0 and b["status"] == 0
    assert a["published"] == b["published"] == 15 and a["header_ok"]
    assert a["n_objects"] == b["n_objects"]
    for k in ("id", "center", "bbox"):
        assert a["objects"][k].tobytes() == b["objects"][k].tobytes(), k
    assert (a["objects_raw"][:, 4:8] == [0, 0, 0, 1]).all()
    va, vb = a["objects"]["velocity"], b["objects"]["velocity"]   # std::sort ties: the selected speed is defined
    na = np.sqrt((va[:, 0] * va[:, 0] + va[:, 1] * va[:, 1]) + va[:, 2] * va[:, 2])
    nb = np.sqrt((vb[:, 0] * vb[:, 0] + vb[:, 1] * vb[:, 1]) + vb[:, 2] * vb[:, 2])
    assert np.array_equal(na, nb)
    assert np.array_equal(a["clusters_image"], b["clusters_image"]), "~clusters_image"
    assert a["n_markers"] == b["n_markers"]
    for k, (ma, mb) in enumerate(zip(a["markers"], b["markers"])):
        assert np.array_equal(ma, mb), f"marker {k}"
    assert np.array_equal(a["marker_rgba"], b["marker_rgba"])
    return a, b


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 4, 10])
def test_gpu_nodelet_publishes_what_the_reference_nodelet_publishes(plugin, n):
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not in the snapshot")
    from moving_object_detector_b200 import synth
    cloud, _ = synth.make_stress_cloud(736, 414, seed=4004, cluster_size=300, fx=466.0)
    a, _ = _compare(plugin, ref_py, cloud, cluster_size=300, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=n)
    assert a["n_objects"] >= 3


@pytest.mark.gpu
def test_gpu_nodelet_on_a_constructor_cloud_and_subscriber_gating(plugin, oracle):
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not in the snapshot")
    from helpers import oracle_scene_flow
    from moving_object_detector_b200 import synth
    seq = synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)
    cloud = oracle_scene_flow(oracle, seq, 0)["points32"]
    a, _ = _compare(plugin, ref_py, cloud)
    assert a["n_objects"] >= 2
    for subs in (0, 1, 2, 4, 5):
        r, q = plugin.cluster(cloud, subscribers=subs), ref_py.cluster(cloud, subscribers=subs)
        assert r["published"] & 7 == q["published"] & 7 == subs


@pytest.mark.gpu
def test_gpu_constructor_host_class_publishes_what_the_reference_node_publishes(plugin):
    """SceneFlowConstructorB200::construct vs SceneFlowConstructor::construct at the ROS message level: payloads bit for
    bit, and headers / layout / encodings / publication order of ~optical_flow, ~depth, ~scene_flow,
    ~synthetic_optical_flow as strings (oracle/msg_describe.h), under the degrade rules and subscriber gates."""
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not in the snapshot")
    from helpers import assert_cloud_bit_exact
    from moving_object_detector_b200 import synth
    seq = synth.make_sequence(synth.CONFIGS["gazebo_640x480"], 2)
    disp, flows = seq["disp"].numpy(), seq["flows"].numpy()
    now, prev, flow = (np.ascontiguousarray(a) for a in (disp[1], disp[0], flows[0]))
    sp, sn, _ = ref_py.stamps_for(seq["dts"][0])
    kw = dict(f=seq["f"], T=seq["T"], min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"])
    tr = seq["transforms"][0]
    fields = [0, 1, 2, 4, 5, 6]
    cases = [(now, prev, flow, tr, 15), (now, prev, flow, tr, 1), (now, prev, flow, tr, 6), (now, prev, flow, tr, 0),
             (now, prev, None, tr, 15), (now, None, flow, tr, 15), (now, prev, flow, None, 15), (None, prev, flow, tr, 15)]
    for a_now, a_prev, a_flow, a_tr, subs in cases:
        r = ref_py.scene_flow(a_now, a_prev, a_flow, seq["cam"], a_tr, sp, sn, subscribers=subs, **kw)
        g = plugin.scene_flow(a_now, a_prev, a_flow, seq["cam"], a_tr, sp, sn, subscribers=subs, **kw)
        assert r["status"] == 0 and g["status"] == 0
        assert g["published"] == r["published"], (subs, g["published"], r["published"])
        assert g["description"] == r["description"], f"message metadata / order differ:\n{g['description']}\n{r['description']}"
        if r["published"] & 1:
            assert_cloud_bit_exact(g["points32"][..., fields], r["points32"][..., fields], "~scene_flow")
        if r["published"] & 2:
            assert_cloud_bit_exact(g["depth"][..., None], r["depth"][..., None], "~depth")
        if r["published"] & 4:
            assert_cloud_bit_exact(g["static_flow"], r["static_flow"], "~synthetic_optical_flow")
        if r["published"] & 8:
            assert g["flow_echo"].tobytes() == r["flow_echo"].tobytes()   # the flow holds NaNs: compare bytes
