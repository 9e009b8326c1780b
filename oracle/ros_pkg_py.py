"""ctypes binding of oracle/_shim/libclusterer_nodelet_b200.so: the PRODUCT's drop-in ROS nodelet
(moving_object_detector_b200/ros/scene_flow_clusterer, written against the real ROS API) compiled against the stand-in
ROS headers of oracle/ref_shim/ and driven by oracle/ros_pkg_wrap.cpp.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle_py import OBJECT_DTYPE

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
    return _lib


def defaults():
    lib = load()
    i3 = (C.c_int32 * 3)()
    d2 = (C.c_double * 2)()
    assert lib.rosb200_defaults(i3, d2) == 0
    return dict(cluster_size=i3[0], neighbor_distance=i3[1], topics_ok=i3[2], depth_diff=d2[0], dynamic_speed=d2[1])


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
