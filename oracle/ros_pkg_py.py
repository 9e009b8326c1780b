"""ctypes binding of oracle/_shim/libclusterer_nodelet_b200.so: the PRODUCT's drop-in ROS nodelet
(moving_object_detector_b200/ros/scene_flow_clusterer, written against the real ROS API) compiled against the stand-in
ROS headers of oracle/ref_shim/ and driven by oracle/ros_pkg_wrap.cpp.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle_py import OBJECT_DTYPE, Transform, _cam, _disp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_shim", "libclusterer_nodelet_b200.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.rosb200_defaults.restype = C.c_int
        _lib.rosb200_cluster.restype = C.c_int
        _lib.rosb200_scene_flow.restype = C.c_int
    return _lib


def defaults():
    lib = load()
    i3 = (C.c_int32 * 3)()
    d2 = (C.c_double * 2)()
    assert lib.rosb200_defaults(i3, d2) == 0
    return dict(cluster_size=i3[0], neighbor_distance=i3[1], topics_ok=i3[2], depth_diff=d2[0], dynamic_speed=d2[1])


def scene_flow(disp_now, disp_prev, flow, cam, transform, stamp_prev, stamp_now, dynamic_flow_diff=5, f=1.0, T=1.0,
               min_disparity=0.0, max_disparity=128.0, subscribers=7):
    """SceneFlowConstructorB200::construct on one pair; arguments and result keys as ref_py.scene_flow."""
    lib = load()
    ref_img = disp_now if disp_now is not None else disp_prev
    h, w = ref_img.shape if ref_img is not None else flow.shape[:2]
    out = np.full((h, w, 8), -12345.0, np.float32)
    depth = np.full((h, w), -12345.0, np.float32)
    static = np.full((h, w, 2), -12345.0, np.float32)
    echo = np.full((h, w, 2), -12345.0, np.float32)
    pub = C.c_int32(0)
    c = _cam(cam)
    dn = _disp(disp_now, f, T, min_disparity, max_disparity) if disp_now is not None else None
    dp = _disp(disp_prev, f, T, min_disparity, max_disparity) if disp_prev is not None else None
    t = Transform(*[float(v) for v in transform]) if transform is not None else None
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    buf = C.create_string_buffer(4096)
    rc = lib.rosb200_scene_flow(w, h, C.byref(c), C.byref(dn) if dn else None, C.byref(dp) if dp else None, p(flow),
                                C.c_int32(flow.strides[0] if flow is not None else 0), C.byref(t) if t else None,
                                C.c_uint32(stamp_prev[0]), C.c_uint32(stamp_prev[1]), C.c_uint32(stamp_now[0]),
                                C.c_uint32(stamp_now[1]), C.c_int32(dynamic_flow_diff), C.c_int(subscribers), p(out),
                                p(depth), p(static), p(echo), C.byref(pub), buf, 4096)
    return dict(status=rc, published=pub.value, points32=out, depth=depth, static_flow=static, flow_echo=echo,
                description=buf.value.decode())


def cluster(points32, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=4, max_objs=4096,
            subscribers=7, max_markers=65536):
    """ClustererNodelet::dataCB of the product's nodelet on one (H,W,8) float32 cloud; same result keys as
    ref_py.cluster (minus the reference's internal maps)."""
    lib = load()
    points32 = np.ascontiguousarray(points32, np.float32)
    h, w = points32.shape[:2]
    objs = np.zeros((max_objs, 14), np.float64)
    n = C.c_int32(0)
    bgr = np.zeros((h, w, 3), np.uint8)
    moff = np.zeros(max_markers + 1, np.int32)
    mpts = np.zeros((h * w, 3), np.float64)
    mrgba = np.zeros((max_markers, 4), np.float32)
    nm, pub, hdr = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    rc = lib.rosb200_cluster(w, h, p(points32), C.c_int32(cluster_size), C.c_double(depth_diff),
                             C.c_double(dynamic_speed), C.c_int32(neighbor_distance), C.c_int(subscribers), p(objs),
                             C.c_int32(max_objs), C.byref(n), p(bgr), p(moff), p(mpts), p(mrgba),
                             C.c_int32(max_markers), C.byref(nm), C.byref(pub), C.byref(hdr))
    k = min(n.value, max_objs)
    raw = objs[:k]
    rec = np.zeros(k, OBJECT_DTYPE)
    rec["id"] = raw[:, 0].astype(np.int32)
    rec["center"] = raw[:, 1:4].astype(np.float32)
    rec["velocity"] = raw[:, 8:11].astype(np.float32)
    rec["bbox"] = raw[:, 11:14].astype(np.float32)
    m = min(nm.value, max_markers)
    return dict(status=rc, published=pub.value, header_ok=bool(hdr.value), objects=rec, objects_raw=raw.copy(),
                n_objects=n.value, clusters_image=bgr, markers=[mpts[moff[i]:moff[i + 1]].copy() for i in range(m)],
                marker_rgba=mrgba[:m].copy(), n_markers=nm.value)


def node_sequence(disps, flows, motions, ok_flags, stamps, cam, f=1.0, T=1.0, min_disparity=0.0, max_disparity=128.0,
                  subscribers=15, dynamic_flow_diff=-1, fused=False, cluster_size=2500):
    """The PRODUCT's constructor node (ros/scene_flow_constructor) through stereoCallback; arguments and result as
    ref_py.node_sequence."""
    from ref_py import _node_sequence
    return _node_sequence(load().rosb200_node_sequence, disps, flows, motions, ok_flags, stamps, cam, f, T,
                          min_disparity, max_disparity, subscribers, dynamic_flow_diff, fused, cluster_size)


def constructor_defaults():
    lib = load()
    i2 = (C.c_int32 * 2)()
    lib.rosb200_constructor_defaults.restype = C.c_int
    assert lib.rosb200_constructor_defaults(i2) == 0
    return dict(dynamic_flow_diff=i2[0], topics_ok=i2[1])
