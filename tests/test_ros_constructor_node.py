"""The drop-in ROS package of the constructor (moving_object_detector_b200/ros/scene_flow_constructor): the product's
SceneFlowConstructor NODE, written against the real ROS API, compiled unmodified against the stand-in ROS headers
(oracle/ref_shim) and driven through stereoCallback frame after frame next to the REFERENCE's node
(oracle/_ref/libmod_ref.so), both fed by the same scripted sgm_gpu / pwc_net / libviso2 stand-ins.  Everything the two
nodes publish (~optical_flow, ~depth, ~scene_flow, ~synthetic_optical_flow) and broadcast on tf must be equal message
for message: topic order, headers, layouts and payloads, across estimator failures and subscriber gates.

CPU: package surface, start-up defaults, loud stop without a GPU, and what the reference node itself does over a
sequence (this pins the harness).  GPU (-m gpu): parity of the two nodes' logs."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "moving_object_detector_b200", "ros", "scene_flow_constructor")
W, H = 320, 240


@pytest.fixture(scope="module")
def plugin():
    import ros_pkg_py
    if not ros_pkg_py.available():
        pytest.skip("oracle/_shim/libclusterer_nodelet_b200.so not built (make -C oracle ros_pkg)")
    ros_pkg_py.load()
    return ros_pkg_py


@pytest.fixture(scope="module")
def ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libmod_ref.so not in the snapshot")
    ref_py.load()
    return ref_py


def _motion(t7):
    """libviso2-style 4x4 matrix (previous camera frame -> current) of a (tx,ty,tz,qx,qy,qz,qw) transform."""
    tx, ty, tz, x, y, z, w = (float(v) for v in t7)
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    m = np.eye(4)
    m[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                 [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                 [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
    m[:3, 3] = [tx, ty, tz]
    return m


@pytest.fixture(scope="module")
def script():
    """Five stereo frames of a small Gazebo-like sequence as the upstream estimators would report them."""
    import dataclasses
    from moving_object_detector_b200 import synth
    base = synth.CONFIGS["gazebo_640x480"]
    cfg = dataclasses.replace(base, w=W, h=H, fx=base.fx / 2, cx=base.cx / 2, cy=base.cy / 2)
    n = 5
    seq = synth.make_sequence(cfg, n)
    disps = np.ascontiguousarray(seq["disp"].numpy(), np.float32)
    flows = np.zeros((n, H, W, 2), np.float32)
    flows[1:] = seq["flows"].numpy()
    motions = np.stack([np.eye(4)] + [_motion(t) for t in seq["transforms"]])
    stamps = np.array([[1000 + (k * 33333333) // 1000000000, (k * 33333333) % 1000000000] for k in range(n)], np.uint32)
    kw = dict(cam=seq["cam"], f=seq["f"], T=seq["T"], min_disparity=seq["min_disparity"], max_disparity=seq["max_disparity"])
    return dict(disps=disps, flows=flows, motions=motions, stamps=stamps, kw=kw, n=n)


def _run(side, script, ok=None, **over):
    ok = np.full(script["n"], 7, np.int32) if ok is None else np.asarray(ok, np.int32)
    kw = dict(script["kw"])
    kw.update(over)
    return side.node_sequence(script["disps"], script["flows"], script["motions"], ok, script["stamps"], **kw)


def _frames(log):
    """{frame index: [topic, ...]} of a node_sequence log."""
    out, cur = {}, None
    for line in log.splitlines():
        m = re.match(r"frame (\d+)$", line)
        if m:
            cur = out.setdefault(int(m.group(1)), [])
        elif not line.startswith("  "):
            cur.append(line.split(":")[0])
    return out


# ---- CPU ------------------------------------------------------------------------------------------------------------
def test_package_surface_matches_the_reference_contract():
    """Package / executable / topic / parameter names of SURVEY.md 8(b) (reference: scene_flow_constructor/package.xml,
    CMakeLists.txt:43-46, cfg/SceneFlowConstructor.cfg:8-9, src/scene_flow_constructor.cpp:27-62)."""
    assert "<name>scene_flow_constructor</name>" in open(os.path.join(PKG, "package.xml")).read()
    cm = open(os.path.join(PKG, "CMakeLists.txt")).read()
    assert "project(scene_flow_constructor)" in cm and "add_executable(scene_flow_constructor " in cm
    assert "generate_dynamic_reconfigure_options(cfg/SceneFlowConstructor.cfg)" in cm
    cfg = open(os.path.join(PKG, "cfg", "SceneFlowConstructor.cfg")).read()
    params = re.findall(r'gen\.add\("(\w+)", (\w+), 0, "[^"]*", ([\d.]+), ([\d.]+), ([\d.]+)\)', cfg)
    assert params == [("dynamic_flow_diff", "int_t", "5", "1", "100"), ("max_color_velocity", "double_t", "1.0", "0.1", "10.0")]
    assert 'gen.generate(PACKAGE, "scene_flow_constructor", "SceneFlowConstructor")' in cfg
    src = open(os.path.join(PKG, "src", "scene_flow_constructor.cpp")).read()
    for needle in ('advertise<sensor_msgs::Image>("depth", 1)', 'advertise<sensor_msgs::Image>("optical_flow", 1)',
                   'advertise<sensor_msgs::PointCloud2>("scene_flow", 1)',
                   'advertise<sensor_msgs::Image>("synthetic_optical_flow", 1)', 'resolveName("left_image")',
                   'resolveName("right_image")', '"visual_odometry"', '"base_link_frame_id"', '"odom_frame_id"'):
        assert needle in src, needle
    assert 'ros::init(argc, argv, "scene_flow_constructor")' in open(os.path.join(PKG, "src", "scene_flow_constructor_node.cpp")).read()


def test_if_the_reference_is_here_its_surface_is_the_same():
    ref_dir = "/root/reference/scene_flow_constructor"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present")
    grab = lambda s: re.findall(r'gen\.add\("(\w+)", (\w+), 0, "[^"]*", ([\d.]+), ([\d.]+), ([\d.]+)\)', s)
    assert grab(open(os.path.join(PKG, "cfg", "SceneFlowConstructor.cfg")).read()) == \
        grab(open(os.path.join(ref_dir, "cfg", "SceneFlowConstructor.cfg")).read())
    # every ~visual_odometry parameter the reference reads is read here under the same name
    theirs = set(re.findall(r'getParam\("(\w+)"', open(os.path.join(ref_dir, "include", "odometry_params.h")).read()))
    ours = set(re.findall(r'getParam\("(\w+)"', open(os.path.join(PKG, "src", "scene_flow_constructor.cpp")).read()))
    assert theirs and theirs <= ours
    # topics
    ref_src = open(os.path.join(ref_dir, "src", "scene_flow_constructor.cpp")).read()
    src = open(os.path.join(PKG, "src", "scene_flow_constructor.cpp")).read()
    adv = lambda s: sorted(re.findall(r'advertise<([\w:]+)>\("(\w+)", (\d+)\)', s))
    # the reference's four topics, plus ~moving_objects of the optional in-process clusterer (~fused_clusterer)
    assert [a for a in adv(src) if a[1] != "moving_objects"] == adv(ref_src)


def test_node_latches_cfg_defaults_and_advertises_the_topics(plugin):
    assert plugin.constructor_defaults() == dict(dynamic_flow_diff=5, topics_ok=4)


def test_node_stops_loudly_without_a_gpu(plugin, script):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(plugin, script)
    assert r["status"] == 0 and r["shutdown_requested"]   # ROS_FATAL + ros::shutdown(): no CPU fallback
    topics = _frames(r["log"])
    assert all(t == ["/tf"] for t in topics.values()), topics   # the odometry is still broadcast; nothing else


def test_reference_node_over_a_sequence(ref, script):
    """What the reference node does frame after frame (pins the harness and the degrade rules the product must follow):
    frame 0 has no previous image -> only tf + ~depth; later frames publish ~optical_flow, ~depth, ~scene_flow,
    ~synthetic_optical_flow in that order; a failed estimator removes exactly what depends on it, on this frame and -
    for the disparity - on the next."""
    r = _run(ref, script)
    assert r["status"] == 0 and not r["shutdown_requested"]
    t = _frames(r["log"])
    full = ["/tf", "optical_flow", "depth", "scene_flow", "synthetic_optical_flow"]
    assert t == {0: ["/tf", "depth"], 1: full, 2: full, 3: full, 4: full}
    assert "frame=left_camera_optical" in r["log"] and "child=base_link" in r["log"] and "frame=odom" in r["log"]
    # ~scene_flow payloads differ from frame to frame (the script really advances)
    payloads = re.findall(r"scene_flow: .*\n  payload=(\w+) finite=(\d+)", r["log"])
    assert len(payloads) == 4 and len(set(p for p, _ in payloads)) == 4
    assert all(int(n) > 0.8 * 6 * W * H for _, n in payloads), payloads   # real clouds, not NaN-filled ones

    # disparity fails on frame 2: no ~depth / ~scene_flow there, and frame 3 has no previous disparity
    t = _frames(_run(ref, script, ok=[7, 7, 6, 7, 7])["log"])
    assert t[2] == ["/tf", "optical_flow"] and t[3] == ["/tf", "optical_flow", "depth"] and t[4] == full
    # odometry fails on frame 3: no tf broadcast, no ~scene_flow; frame 4 is complete again
    t = _frames(_run(ref, script, ok=[7, 7, 7, 5, 7])["log"])
    assert t[3] == ["optical_flow", "depth"] and t[4] == full
    # optical flow fails on frame 1: only ~depth
    t = _frames(_run(ref, script, ok=[7, 3, 7, 7, 7])["log"])
    assert t[1] == ["/tf", "depth"] and t[2] == full
    # subscriber gates
    t = _frames(_run(ref, script, subscribers=1)["log"])
    assert t[0] == ["/tf"] and t[1] == ["/tf", "scene_flow"]


# ---- GPU: the two nodes side by side --------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("ok,subs,dfd", [
    (None, 15, -1),                 # everything works, every topic subscribed, default dynamic_flow_diff
    (None, 1, -1), (None, 2, -1), (None, 12, -1), (None, 0, -1),
    ([7, 7, 6, 7, 7], 15, -1),      # sgm_gpu fails on frame 2
    ([7, 7, 7, 5, 7], 15, -1),      # libviso2 fails on frame 3
    ([7, 3, 7, 3, 7], 15, -1),      # pwc_net fails on frames 1 and 3
    ([6, 5, 3, 7, 7], 15, -1),      # a bad start: each estimator fails once
    (None, 15, 2), (None, 15, 40),  # dynamic_reconfigure values
])
def test_gpu_node_publishes_what_the_reference_node_publishes(plugin, ref, script, ok, subs, dfd):
    a = _run(plugin, script, ok=ok, subscribers=subs, dynamic_flow_diff=dfd)
    b = _run(ref, script, ok=ok, subscribers=subs, dynamic_flow_diff=dfd)
    assert a["status"] == 0 and b["status"] == 0 and not a["shutdown_requested"]
    assert a["log"] == b["log"], "the nodes' publications differ:\n--- product\n%s\n--- reference\n%s" % (a["log"], b["log"])
    if subs == 15 and ok is None:
        assert a["log"].count("scene_flow: ") == script["n"] - 1


# ---- ~fused_clusterer: the clusterer in the constructor's process -----------------------------------------------------
def _objects_blocks(log):
    """The moving_objects messages of a log, one string per message (header line + the objects' lines)."""
    out, cur = [], None
    for line in log.splitlines():
        if line.startswith("moving_objects: "):
            cur = [line]
            out.append(cur)
        elif cur is not None and line.startswith("  "):
            cur.append(line)
        else:
            cur = None
    return ["\n".join(b) for b in out]


def test_reference_two_node_chain_over_a_sequence(ref, script):
    """constructor -> ~scene_flow -> clusterer nodelet -> ~moving_objects with the reference's own nodes: what the
    product's ~fused_clusterer option has to reproduce (and a check of the harness)."""
    r = _run(ref, script, subscribers=1, fused=True, cluster_size=300)
    assert r["status"] == 0
    t = _frames(r["log"])
    assert t[0] == ["/tf"] and all(t[k] == ["/tf", "scene_flow", "moving_objects"] for k in range(1, script["n"]))
    blocks = _objects_blocks(r["log"])
    assert len(blocks) == script["n"] - 1 and all("objects=2" in b for b in blocks)
    assert "frame=left_camera_optical" in blocks[0]   # the header of the cloud, i.e. of the left image


def test_fused_option_is_part_of_the_package():
    src = open(os.path.join(PKG, "src", "scene_flow_constructor.cpp")).read()
    assert 'param("fused_clusterer", fused_clusterer_, false)' in src
    assert 'advertise<moving_object_msgs::MovingObjectArray>("moving_objects", 1)' in src
    assert 'ros::NodeHandle(pnh, "clusterer")' in src


@pytest.mark.gpu
@pytest.mark.parametrize("ok", [None, [7, 7, 6, 7, 7], [7, 3, 7, 5, 7]])
def test_gpu_fused_node_publishes_what_the_reference_two_node_chain_publishes(plugin, ref, script, ok):
    want = _run(ref, script, ok=ok, subscribers=1, fused=True, cluster_size=300)
    got = _run(plugin, script, ok=ok, subscribers=1, fused=True, cluster_size=300)
    assert want["status"] == 0 and got["status"] == 0 and not got["shutdown_requested"]
    assert got["log"] == want["log"], "publications differ:\n--- product\n%s\n--- reference\n%s" % (got["log"], want["log"])
    # nobody subscribes to ~scene_flow: the cloud never leaves the GPU, the objects are the same
    quiet = _run(plugin, script, ok=ok, subscribers=0, fused=True, cluster_size=300)
    assert quiet["status"] == 0 and "scene_flow: " not in quiet["log"]
    assert _objects_blocks(quiet["log"]) == _objects_blocks(want["log"])
    if ok is None:
        assert len(_objects_blocks(got["log"])) == script["n"] - 1
