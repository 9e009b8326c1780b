"""ctypes binding of oracle/_ref/libmod_ref.so: the REFERENCE's own hot-path translation units
(disparity_image_processor.cpp, scene_flow_constructor.cpp, clusterer_nodelet.cpp, color_set.cpp, lookup_table.cpp),
compiled unmodified from /root/reference against the stand-in dependency headers of oracle/ref_shim/ (see
oracle/Makefile target ref_full and oracle/ref_full_wrap.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.  The library is built only where
/root/reference exists (this container); the built .so travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle_py import Camera, Disparity, Transform, _cam, _disp, OBJECT_DTYPE, ENCODINGS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmod_ref.so")
COLORMAP_PATH = os.path.join(os.path.dirname(_HERE), "tests", "golden", "hsv_colormap_bgr.npy")

SUB_SCENE_FLOW, SUB_DEPTH, SUB_STATIC_FLOW, SUB_OPTICAL_FLOW = 1, 2, 4, 8
SUB_MOVING_OBJECTS, SUB_CLUSTERS, SUB_CLUSTERS_IMAGE = 1, 2, 4

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        for fn in ("ref_set_colormap", "ref_defaults", "ref_scene_flow", "ref_cluster", "ref_bench_pipeline",
                   "ref_image_crop"):
            getattr(_lib, fn).restype = C.c_int
        cm = np.ascontiguousarray(np.load(COLORMAP_PATH), np.uint8)
        assert cm.shape == (256, 3)
        _lib.ref_set_colormap(C.c_void_p(cm.ctypes.data))
    return _lib


def stamps_for(dt_sec, base_sec=1000):
    """ROS stamps (prev, now) whose difference is dt rounded to the nanosecond, and the double that
    (stamp_now - stamp_prev).toSec() yields for them: (double)sec + 1e-9 * (double)nsec."""
    ns = int(round(float(dt_sec) * 1e9))
    sec, nsec = divmod(ns, 1000000000)
    dt = float(sec) + 1e-9 * float(nsec)
    return (base_sec, 0), (base_sec + sec, nsec), dt


def defaults():
    lib = load()
    i3 = (C.c_int32 * 3)()
    d2 = (C.c_double * 2)()
    rc = lib.ref_defaults(i3, d2)
    assert rc == 0, rc
    return dict(dynamic_flow_diff=i3[0], cluster_size=i3[1], neighbor_distance=i3[2], depth_diff=d2[0],
                dynamic_speed=d2[1])


def scene_flow(disp_now, disp_prev, flow, cam, transform, stamp_prev, stamp_now, dynamic_flow_diff=5, f=1.0, T=1.0,
               min_disparity=0.0, max_disparity=128.0, subscribers=SUB_SCENE_FLOW | SUB_DEPTH | SUB_STATIC_FLOW):
    """SceneFlowConstructor::construct on one pair.  Any of disp_now / disp_prev / flow / transform may be None.
    Returns dict(status, published, points32, depth, static_flow, flow_echo, cloud_meta)."""
    lib = load()
    ref_img = disp_now if disp_now is not None else disp_prev
    if ref_img is not None:
        h, w = ref_img.shape
    else:
        h, w = flow.shape[:2]
    out = np.full((h, w, 8), -12345.0, np.float32)
    depth = np.full((h, w), -12345.0, np.float32)
    static = np.full((h, w, 2), -12345.0, np.float32)
    echo = np.full((h, w, 2), -12345.0, np.float32)
    meta = np.zeros(20, np.int32)
    pub = C.c_int32(0)
    c = _cam(cam)
    dn = _disp(disp_now, f, T, min_disparity, max_disparity) if disp_now is not None else None
    dp = _disp(disp_prev, f, T, min_disparity, max_disparity) if disp_prev is not None else None
    t = Transform(*[float(v) for v in transform]) if transform is not None else None
    if flow is not None:
        assert flow.dtype == np.float32 and flow.shape == (h, w, 2) and flow.strides[1] == 8
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    rc = lib.ref_scene_flow(w, h, C.byref(c), C.byref(dn) if dn else None, C.byref(dp) if dp else None, p(flow),
                            C.c_int32(flow.strides[0] if flow is not None else 0), C.byref(t) if t else None,
                            C.c_uint32(stamp_prev[0]), C.c_uint32(stamp_prev[1]), C.c_uint32(stamp_now[0]),
                            C.c_uint32(stamp_now[1]), C.c_int32(dynamic_flow_diff), C.c_int(subscribers), p(out),
                            p(depth), p(static), p(echo), C.byref(pub), p(meta))
    buf = C.create_string_buffer(4096)
    lib.ref_last_description(buf, 4096)
    return dict(status=rc, published=pub.value, points32=out, depth=depth, static_flow=static, flow_echo=echo,
                cloud_meta=meta, description=buf.value.decode())


def cluster(points32, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=4, max_objs=4096,
            subscribers=SUB_MOVING_OBJECTS | SUB_CLUSTERS | SUB_CLUSTERS_IMAGE, shape=None, max_markers=65536):
    """ClustererNodelet::dataCB on one cloud ((H,W,8) float32), or on the PointCloud2 the last scene_flow() call
    published when points32 is None (shape=(h, w) then required)."""
    lib = load()
    if points32 is not None:
        points32 = np.ascontiguousarray(points32, np.float32)
        h, w = points32.shape[:2]
    else:
        h, w = shape
    dyn = np.empty((h, w), np.uint8)
    labels = np.empty((h, w), np.int32)
    nclusters = C.c_int32(0)
    objs = np.zeros((max_objs, 14), np.float64)
    n = C.c_int32(0)
    bgr = np.zeros((h, w, 3), np.uint8)
    moff = np.zeros(max_markers + 1, np.int32)
    mpts = np.zeros((h * w, 3), np.float64)
    mrgba = np.zeros((max_markers, 4), np.float32)
    nm = C.c_int32(0)
    pub = C.c_int32(0)
    p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None
    rc = lib.ref_cluster(w, h, p(points32), C.c_int32(cluster_size), C.c_double(depth_diff), C.c_double(dynamic_speed),
                         C.c_int32(neighbor_distance), C.c_int(subscribers), p(dyn), p(labels), C.byref(nclusters),
                         p(objs), C.c_int32(max_objs), C.byref(n), p(bgr), p(moff), p(mpts), p(mrgba),
                         C.c_int32(max_markers), C.byref(nm), C.byref(pub))
    k = min(n.value, max_objs)
    raw = objs[:k]
    # the message carries doubles widened from the reference's float arithmetic; narrowing back is lossless
    rec = np.zeros(k, OBJECT_DTYPE)
    rec["id"] = raw[:, 0].astype(np.int32)
    rec["center"] = raw[:, 1:4].astype(np.float32)
    rec["velocity"] = raw[:, 8:11].astype(np.float32)
    rec["bbox"] = raw[:, 11:14].astype(np.float32)
    lossless = bool(np.array_equal(rec["center"].astype(np.float64), raw[:, 1:4]) and
                    np.array_equal(rec["velocity"].astype(np.float64), raw[:, 8:11]) and
                    np.array_equal(rec["bbox"].astype(np.float64), raw[:, 11:14]))
    m = min(nm.value, max_markers)
    markers = [mpts[moff[i]:moff[i + 1]].copy() for i in range(m)]
    return dict(status=rc, published=pub.value, dynamic=dyn, labels=labels, number_of_clusters=nclusters.value,
                objects=rec, objects_raw=raw.copy(), objects_lossless=lossless, n_objects=n.value, clusters_image=bgr,
                markers=markers, marker_rgba=mrgba[:m].copy(), n_markers=nm.value)


def image_crop(image, encoding, target_w, target_h, K, P):
    """image_crop::ImageCrop::cameraCallback on one image + CameraInfo.  Returns dict(status, image, K, P, width, height,
    image_wh_step); status -12 = the reference threw (crop window outside the image)."""
    lib = load()
    assert image.dtype == np.uint8 and image.strides[-1] == 1
    h, w = image.shape[:2]
    out = np.zeros((max(target_h, 0), max(target_w, 0), 3), np.uint8)
    K = np.ascontiguousarray(K, np.float64).ravel()
    P = np.ascontiguousarray(P, np.float64).ravel()
    Ko, Po, wh = np.zeros(9), np.zeros(12), np.zeros(5, np.int32)
    row_bytes = image[0].size if h else 0
    step = image.strides[0] if h > 1 else row_bytes
    buf = np.zeros((h, step), np.uint8)   # the message's data: h rows of `step` bytes (padding zeroed)
    buf[:, :row_bytes] = image.reshape(h, row_bytes)
    rc = lib.ref_image_crop(w, h, ENCODINGS[encoding], C.c_void_p(buf.ctypes.data), C.c_int32(step),
                            C.c_void_p(K.ctypes.data), C.c_void_p(P.ctypes.data), C.c_int(target_w), C.c_int(target_h),
                            C.c_void_p(out.ctypes.data), C.c_void_p(Ko.ctypes.data), C.c_void_p(Po.ctypes.data),
                            C.c_void_p(wh.ctypes.data))
    return dict(status=rc, image=out, K=Ko, P=Po, width=int(wh[0]), height=int(wh[1]), image_wh_step=tuple(int(v) for v in wh[2:5]))


def bench_pipeline(disp_frames, prev_index, now_index, flows, cam, transforms, dts, f, T, min_disparity,
                   max_disparity, dynamic_flow_diff=5, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3,
                   neighbor_distance=4, n_threads=1, reps=1, stage_mask=3):
    """Times the reference's own construct() + dataCB() over len(transforms) pairs; (best_seconds, n_objects)."""
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
    rc = lib.ref_bench_pipeline(
        w, h, C.byref(c), C.c_float(f), C.c_float(T), C.c_float(min_disparity), C.c_float(max_disparity),
        C.c_void_p(disp_frames.ctypes.data), C.c_void_p(prev_index.ctypes.data), C.c_void_p(now_index.ctypes.data),
        C.c_void_p(flows.ctypes.data), C.cast(tr, C.c_void_p), C.c_void_p(dt.ctypes.data), C.c_int(n_pairs),
        C.c_int32(dynamic_flow_diff), C.c_int32(cluster_size), C.c_double(depth_diff), C.c_double(dynamic_speed),
        C.c_int32(neighbor_distance), C.c_int(n_threads), C.c_int(reps), C.c_int(stage_mask), C.byref(best),
        C.byref(nobj))
    assert rc == 0, rc
    return best.value, nobj.value


def _node_sequence(fn, disps, flows, motions, ok_flags, stamps, cam, f, T, min_disparity, max_disparity, subscribers,
                   dynamic_flow_diff, fused=False, cluster_size=2500, log_cap=1 << 16):
    """Shared marshalling of ref_node_sequence / rosb200_node_sequence (oracle/node_sequence.h)."""
    disps = np.ascontiguousarray(disps, np.float32)
    n, h, w = disps.shape
    flows = np.ascontiguousarray(flows, np.float32)
    assert flows.shape == (n, h, w, 2)
    motions = np.ascontiguousarray(motions, np.float64)
    assert motions.shape == (n, 4, 4)
    ok = np.ascontiguousarray(ok_flags, np.int32)
    st = np.ascontiguousarray(stamps, np.uint32)
    assert ok.shape == (n,) and st.shape == (n, 2)
    c = _cam(cam)
    buf = C.create_string_buffer(log_cap)
    shut = C.c_int32(0)
    p = lambda a: C.c_void_p(a.ctypes.data)
    fn.restype = C.c_int
    rc = fn(w, h, n, C.byref(c), C.c_float(f), C.c_float(T), C.c_float(min_disparity), C.c_float(max_disparity),
            p(disps), p(flows), p(motions), p(ok), p(st), C.c_int(subscribers), C.c_int32(dynamic_flow_diff),
            C.c_int(1 if fused else 0), C.c_int32(cluster_size), buf, C.c_int(log_cap), C.byref(shut))
    return dict(status=rc, log=buf.value.decode(), shutdown_requested=bool(shut.value))


def node_sequence(disps, flows, motions, ok_flags, stamps, cam, f=1.0, T=1.0, min_disparity=0.0, max_disparity=128.0,
                  subscribers=15, dynamic_flow_diff=-1, fused=False, cluster_size=2500):
    """fused: every ~scene_flow cloud is also handed to the reference's clusterer nodelet (the two-node chain).
    The reference NODE through SceneFlowConstructor::stereoCallback, one call per frame, with scripted upstream
    estimators.  disps (n,H,W); flows (n,H,W,2), entry k = optical flow frame k-1 -> k (entry 0 unused); motions
    (n,4,4) libviso2 matrices previous -> current; ok_flags (n) bit0 disparity / bit1 motion / bit2 flow succeed;
    stamps (n,2) sec, nsec.  Returns dict(status, log, shutdown_requested); log = every publication and tf broadcast."""
    return _node_sequence(load().ref_node_sequence, disps, flows, motions, ok_flags, stamps, cam, f, T, min_disparity,
                          max_disparity, subscribers, dynamic_flow_diff, fused, cluster_size)
