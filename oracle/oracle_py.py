"""ctypes binding of the CPU oracle (oracle/libmod_oracle.so) and of the compiled reference piece (oracle/_ref).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmod_oracle.so")
REF_LIB_PATH = os.path.join(_HERE, "_ref", "liblookup_table_ref.so")

OBJECT_DTYPE = np.dtype([("id", "<i4"), ("center", "<f4", 3), ("velocity", "<f4", 3), ("bbox", "<f4", 3),
                         ("n_points", "<i4")])


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fx", "fy", "cx", "cy", "Tx", "Ty")]


class Disparity(C.Structure):
    _fields_ = [("f", C.c_float), ("T", C.c_float), ("min_disparity", C.c_float), ("max_disparity", C.c_float),
                ("data", C.c_void_p), ("step_bytes", C.c_int32)]


class Transform(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("tx", "ty", "tz", "qx", "qy", "qz", "qw")]


_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_scene_flow.restype = C.c_int
        _lib.orc_scene_flow_near.restype = C.c_int
        _lib.orc_cluster.restype = C.c_int
        _lib.orc_lookup_table_run.restype = C.c_int
        _lib.orc_bench_pipeline.restype = C.c_int
    return _lib


def _cam(cam):
    return Camera(*[float(v) for v in cam])


def _disp(img, f, T, dmin, dmax):
    assert img.dtype == np.float32 and img.ndim == 2 and img.strides[1] == 4
    return Disparity(f, T, dmin, dmax, img.ctypes.data, img.strides[0])


def scene_flow(disp_now, disp_prev, flow, cam, transform, dt, dynamic_flow_diff=5, f=1.0, T=1.0, min_disparity=0.0,
               max_disparity=128.0, want_depth=False, want_static=False, want_class=False, want_near=False):
    """cam = (fx, fy, cx, cy, Tx, Ty); transform = (tx, ty, tz, qx, qy, qz, qw).  Returns a dict of arrays."""
    lib = load()
    h, w = disp_now.shape
    assert flow.dtype == np.float32 and flow.shape == (h, w, 2) and flow.strides[1] == 8
    out = np.empty((h, w, 8), np.float32)
    depth = np.empty((h, w), np.float32) if want_depth else None
    static = np.empty((h, w, 2), np.float32) if want_static else None
    cls = np.empty((h, w), np.uint8) if want_class else None
    c, t = _cam(cam), Transform(*[float(v) for v in transform])
    dn, dp = _disp(disp_now, f, T, min_disparity, max_disparity), _disp(disp_prev, f, T, min_disparity, max_disparity)
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    rc = lib.orc_scene_flow(w, h, C.byref(c), C.byref(dn), C.byref(dp), p(flow), C.c_int32(flow.strides[0]),
                            C.byref(t), C.c_double(dt), C.c_int32(dynamic_flow_diff), p(out), p(depth), p(static),
                            p(cls))
    assert rc == 0, rc
    res = {"points32": out, "depth": depth, "static_flow": static, "class": cls}
    if want_near:
        near = np.empty((h, w), np.uint8)
        rc = lib.orc_scene_flow_near(w, h, C.byref(c), C.byref(dn), C.byref(dp), p(flow), C.c_int32(flow.strides[0]),
                                     C.byref(t), C.c_int32(dynamic_flow_diff), p(near))
        assert rc == 0, rc
        res["near"] = near
    return res


def cluster(points32, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=4, max_objs=4096,
            want_near=False):
    lib = load()
    points32 = np.ascontiguousarray(points32, np.float32)
    h, w = points32.shape[:2]
    dyn = np.empty((h, w), np.uint8)
    labels = np.empty((h, w), np.int32)
    near = np.empty((h, w), np.uint8) if want_near else None
    objs = np.zeros(max_objs, OBJECT_DTYPE)
    n = C.c_int32(0)
    rc = lib.orc_cluster(w, h, C.c_void_p(points32.ctypes.data), C.c_int32(cluster_size), C.c_double(depth_diff),
                         C.c_double(dynamic_speed), C.c_int32(neighbor_distance), C.c_void_p(dyn.ctypes.data),
                         C.c_void_p(labels.ctypes.data), C.c_void_p(near.ctypes.data) if want_near else None,
                         C.c_void_p(objs.ctypes.data), C.c_int32(max_objs), C.byref(n))
    return {"status": rc, "dynamic": dyn, "labels": labels, "near": near, "objects": objs[:min(n.value, max_objs)],
            "n_objects": n.value}


def clusters_image(labels, number_of_clusters, colormap_bgr):
    """publishClustersImage on a final label map; colormap_bgr: (256,3) uint8 COLORMAP_HSV table."""
    lib = load()
    labels = np.ascontiguousarray(labels, np.int32)
    cm = np.ascontiguousarray(colormap_bgr, np.uint8)
    h, w = labels.shape
    out = np.empty((h, w, 3), np.uint8)
    rc = lib.orc_clusters_image(w, h, C.c_void_p(labels.ctypes.data), C.c_int32(number_of_clusters),
                                C.c_void_p(cm.ctypes.data), C.c_void_p(out.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"orc_clusters_image: {rc}")
    return out


def cluster_members(labels, number_of_clusters):
    """clusterMap2IndicesCluster order (u outer / v inner): list of index arrays, one per cluster."""
    lib = load()
    labels = np.ascontiguousarray(labels, np.int32)
    h, w = labels.shape
    off = np.empty(number_of_clusters + 1, np.int32)
    idx = np.empty(max(int((labels >= 0).sum()), 1), np.int32)
    rc = lib.orc_cluster_members(w, h, C.c_void_p(labels.ctypes.data), C.c_int32(number_of_clusters),
                                 C.c_void_p(off.ctypes.data), C.c_void_p(idx.ctypes.data))
    if rc != 0:
        raise RuntimeError(f"orc_cluster_members: {rc}")
    return [idx[off[c]:off[c + 1]].copy() for c in range(number_of_clusters)]


ENCODINGS = {"bgr8": 0, "rgb8": 1, "bgra8": 2, "rgba8": 3, "mono8": 4}


def image_crop(image, encoding, target_w, target_h):
    """ImageCrop::cameraCallback's image part: (H,W[,C]) uint8 in `encoding` -> (target_h,target_w,3) bgr8, or None where
    the reference's ROI assertion throws."""
    lib = load()
    assert image.dtype == np.uint8 and image.strides[-1] == 1
    h, w = image.shape[:2]
    out = np.empty((max(target_h, 0), max(target_w, 0), 3), np.uint8)
    lib.orc_image_crop.restype = C.c_int
    rc = lib.orc_image_crop(w, h, ENCODINGS[encoding], C.c_void_p(image.ctypes.data), C.c_int32(image.strides[0]),
                            C.c_int(target_w), C.c_int(target_h), C.c_void_p(out.ctypes.data))
    if rc == -12:
        return None
    assert rc == 0, rc
    return out


def to_mono8(image, encoding, rule=0):
    """cv_bridge::toCvCopy(image, MONO8): (H,W[,C]) uint8 in `encoding` -> (H,W) uint8.  rule 0 = OpenCV 3.x constants
    (the reference's), 1 = OpenCV 4.x."""
    lib = load()
    assert image.dtype == np.uint8 and image.strides[-1] == 1
    h, w = image.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib.orc_to_mono8.restype = C.c_int
    rc = lib.orc_to_mono8(w, h, ENCODINGS[encoding], C.c_void_p(image.ctypes.data),
                          C.c_int32(image.strides[0] if h > 1 else image[0].size), C.c_int(rule),
                          C.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    return out


def crop_camera_info(w, h, target_w, target_h, K, P):
    lib = load()
    K = np.ascontiguousarray(K, np.float64).ravel()
    P = np.ascontiguousarray(P, np.float64).ravel()
    Ko, Po = np.empty(9), np.empty(12)
    lib.orc_crop_camera_info.restype = None
    lib.orc_crop_camera_info(C.c_uint32(w), C.c_uint32(h), C.c_int(target_w), C.c_int(target_h),
                             C.c_void_p(K.ctypes.data), C.c_void_p(P.ctypes.data), C.c_void_p(Ko.ctypes.data),
                             C.c_void_p(Po.ctypes.data))
    return Ko, Po


def quat_to_matrix(transform):
    lib = load()
    t = Transform(*[float(v) for v in transform])
    R = np.empty(9, np.float64)
    lib.orc_quat_to_matrix(C.byref(t), C.c_void_p(R.ctypes.data))
    return R.reshape(3, 3)


def lookup_table_run(size, ops, ref=False):
    """ops: (n,3) int32 of (op, a, b); op 0 addLabel, 1 link(a,b), 2 lookup(a).  ref=True runs the REFERENCE's
    own LookupTable compiled from /root/reference into oracle/_ref."""
    ops = np.ascontiguousarray(ops, np.int32)
    out = np.empty(len(ops), np.int32)
    if ref:
        lib = C.CDLL(REF_LIB_PATH)
        fn = lib.ref_lookup_table_run
    else:
        fn = load().orc_lookup_table_run
    fn.restype = C.c_int
    n = fn(C.c_int(size), C.c_void_p(ops.ctypes.data), C.c_int(len(ops)), C.c_void_p(out.ctypes.data),
           C.c_int(len(out)))
    assert n >= 0
    return out[:n].copy()


def bench_pipeline(disp_frames, prev_index, now_index, flows, cam, transforms, dts, f, T, min_disparity,
                   max_disparity, dynamic_flow_diff=5, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3,
                   neighbor_distance=4, n_threads=1, reps=1, stage_mask=3):
    """Times the reference-structured CPU path over len(transforms) pairs; returns (best_seconds, n_objects)."""
    lib = load()
    disp_frames = np.ascontiguousarray(disp_frames, np.float32)
    flows = np.ascontiguousarray(flows, np.float32)
    n_pairs, h, w = flows.shape[0], flows.shape[1], flows.shape[2]
    prev_index = np.ascontiguousarray(prev_index, np.int32)
    now_index = np.ascontiguousarray(now_index, np.int32)
    tr = (Transform * n_pairs)(*[Transform(*[float(v) for v in t]) for t in transforms])
    dt = np.ascontiguousarray(dts, np.float64)
    c = _cam(cam)
    best = C.c_double(0)
    nobj = C.c_int64(0)
    rc = lib.orc_bench_pipeline(
        w, h, C.byref(c), C.c_float(f), C.c_float(T), C.c_float(min_disparity), C.c_float(max_disparity),
        C.c_void_p(disp_frames.ctypes.data), C.c_void_p(prev_index.ctypes.data), C.c_void_p(now_index.ctypes.data),
        C.c_void_p(flows.ctypes.data), C.cast(tr, C.c_void_p), C.c_void_p(dt.ctypes.data), C.c_int(n_pairs),
        C.c_int32(dynamic_flow_diff), C.c_int32(cluster_size), C.c_double(depth_diff), C.c_double(dynamic_speed),
        C.c_int32(neighbor_distance), C.c_int(n_threads), C.c_int(reps), C.c_int(stage_mask), C.byref(best),
        C.byref(nobj))
    assert rc == 0, rc
    return best.value, nobj.value
