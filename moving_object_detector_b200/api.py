"""Python host-side harness above the C ABI (include/mod_b200.h).

This is the test / benchmark harness, not a second implementation: every method marshals numpy arrays or torch
device tensors into one C-ABI call.  Message-like containers mirror what the reference's seams consume
(stereo_msgs/DisparityImage + CameraInfo, geometry_msgs/Transform, the dynamic_reconfigure configs,
moving_object_msgs/MovingObject); the C++ mirror of the node / nodelet classes lives in host/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (Camera, ClParams, Disparity, FrameStats, ModError, Object, PairDev, SfParams, Transform)

POINT_STEP = 32  # sizeof(pcl::PointXYZVelocity), pcl_point_xyz_velocity.h:8-24
OBJECT_DTYPE = np.dtype([("id", "<i4"), ("center", "<f4", 3), ("velocity", "<f4", 3), ("bbox", "<f4", 3),
                         ("n_points", "<i4")])
STATS_DTYPE = np.dtype([("n_dynamic", "<i4"), ("n_clusters", "<i4"), ("n_objects", "<i4"), ("status", "<i4")])
assert OBJECT_DTYPE.itemsize == C.sizeof(Object)
assert STATS_DTYPE.itemsize == C.sizeof(FrameStats)


@dataclass
class CameraInfo:
    """Left sensor_msgs/CameraInfo: only the 3x4 projection matrix P is read (image_geometry fromCameraInfo)."""
    P: Sequence[float]

    def to_c(self) -> Camera:
        P = self.P
        return Camera(fx=P[0], fy=P[5], cx=P[2], cy=P[6], Tx=P[3], Ty=P[7])

    @staticmethod
    def pinhole(fx: float, fy: float, cx: float, cy: float, Tx: float = 0.0, Ty: float = 0.0) -> "CameraInfo":
        return CameraInfo([fx, 0.0, cx, Tx, 0.0, fy, cy, Ty, 0.0, 0.0, 1.0, 0.0])


@dataclass
class DisparityImage:
    """stereo_msgs/DisparityImage as consumed at disparity_image_processor.cpp:3-15."""
    image: np.ndarray  # (H, W) float32; row stride may exceed W*4 (image.step)
    f: float
    T: float
    min_disparity: float = 0.0
    max_disparity: float = 128.0

    def to_c(self) -> Disparity:
        img = self.image
        assert img.dtype == np.float32 and img.ndim == 2 and img.strides[1] == 4
        return Disparity(f=self.f, T=self.T, min_disparity=self.min_disparity, max_disparity=self.max_disparity,
                         data=img.ctypes.data, step_bytes=img.strides[0])


@dataclass
class TransformMsg:
    """geometry_msgs/Transform (scene_flow_constructor.cpp:248-249)."""
    translation: Sequence[float] = (0.0, 0.0, 0.0)
    rotation: Sequence[float] = (0.0, 0.0, 0.0, 1.0)  # x, y, z, w

    def to_c(self) -> Transform:
        t, q = self.translation, self.rotation
        return Transform(tx=t[0], ty=t[1], tz=t[2], qx=q[0], qy=q[1], qz=q[2], qw=q[3])


@dataclass
class SceneFlowConstructorConfig:
    """cfg/SceneFlowConstructor.cfg:8-9"""
    dynamic_flow_diff: int = 5
    max_color_velocity: float = 1.0

    def to_c(self) -> SfParams:
        return SfParams(self.dynamic_flow_diff, self.max_color_velocity)


@dataclass
class ClustererConfig:
    """cfg/Clusterer.cfg:8-11"""
    cluster_size: int = 2500
    depth_diff: float = 0.15
    dynamic_speed: float = 0.3
    neighbor_distance: int = 4

    def to_c(self) -> ClParams:
        return ClParams(self.cluster_size, self.depth_diff, self.dynamic_speed, self.neighbor_distance)


@dataclass
class MovingObject:
    """moving_object_msgs/MovingObject; orientation is always (0,0,0,1) (clusterer_nodelet.cpp:163-166)."""
    id: int
    center: np.ndarray
    velocity: np.ndarray
    bounding_box: np.ndarray
    n_points: int


ENCODINGS = {"bgr8": 0, "rgb8": 1, "bgra8": 2, "rgba8": 3, "mono8": 4}  # MOD_ENC_* of include/mod_b200.h


def crop_camera_info(w: int, h: int, target_w: int, target_h: int, K, P):
    """The CameraInfo fix-up of ImageCrop::cameraCallback (image_crop/src/image_crop.cpp:32-41): returns (K', P');
    the caller sets width / height to the target size.  Raises ModError for a window outside the image."""
    lib = _lib.load()
    K = np.ascontiguousarray(K, np.float64).ravel()
    P = np.ascontiguousarray(P, np.float64).ravel()
    assert K.size == 9 and P.size == 12
    Ko, Po = np.empty(9), np.empty(12)
    rc = lib.mod_host_crop_camera_info(int(w), int(h), int(target_w), int(target_h), K.ctypes.data, P.ctypes.data,
                                       Ko.ctypes.data, Po.ctypes.data)
    if rc != _lib.MOD_OK:
        raise ModError(rc, "crop window does not lie inside the image")
    return Ko, Po


def hsv_colormap() -> np.ndarray:
    """The COLORMAP_HSV table the library embeds, (256, 3) uint8 BGR."""
    out = np.empty((256, 3), np.uint8)
    _lib.load().mod_host_hsv_colormap(out.ctypes.data)
    return out


def objects_from_records(rec: np.ndarray) -> List[MovingObject]:
    return [MovingObject(int(r["id"]), r["center"].astype(np.float64), r["velocity"].astype(np.float64),
                         r["bbox"].astype(np.float64), int(r["n_points"])) for r in rec]


@dataclass
class ClusterResult:
    objects: np.ndarray  # OBJECT_DTYPE records in id order
    labels: Optional[np.ndarray] = None  # (H, W) int32, -1 = NOT_BELONGED_
    points32: Optional[np.ndarray] = None  # (H, W, 8) float32


def _dev_ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


class Context:
    """Owns one mod_ctx (device workspaces + streams).  One call at a time per context."""

    def __init__(self, max_w: int, max_h: int, max_batch: int = 1, device: int = 0, max_clusters: int = 0):
        self._lib = _lib.load()
        handle = C.c_void_p()
        rc = self._lib.mod_create(device, max_w, max_h, max_batch, max_clusters, C.byref(handle))
        if rc != _lib.MOD_OK:
            raise ModError(rc, "mod_create failed (no CUDA device of compute capability 10.x?)")
        self._h = handle
        self.max_w, self.max_h, self.max_batch, self.device = max_w, max_h, max_batch, device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mod_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launch_count(self) -> int:
        return int(self._lib.mod_launch_count(self._h))

    @property
    def stream(self) -> int:
        """The context's own cudaStream_t as an integer handle."""
        return int(self._lib.mod_stream(self._h) or 0)

    @staticmethod
    def set_host_copy_threads(n: int) -> None:
        """Host threads that stage pageable message buffers through pinned memory (process-wide; 0 = driver path)."""
        _lib.load().mod_set_host_copy_threads(int(n))

    @staticmethod
    def host_copy_threads() -> int:
        return _lib.load().mod_host_copy_threads()

    def set_frame_ids(self, now_id: int, prev_id: int) -> None:
        """Identities of the next scene_flow / pipeline call's disparity images (e.g. stamps in ns): a `prev` that was the
        previous call's `now` is then reused on the device instead of uploaded again."""
        self._check(self._lib.mod_set_frame_ids(self._h, int(now_id), int(prev_id)))

    @property
    def resident_hits(self) -> int:
        return int(self._lib.mod_resident_hits(self._h))

    def set_static_flow_filter(self, enabled: bool) -> None:
        """Switches the filtered residual gate of the scene-flow kernel (outputs are identical either way)."""
        self._check(self._lib.mod_set_static_flow_filter(self._h, 1 if enabled else 0))

    def filter_fallbacks(self, reset: bool = True) -> int:
        """Pixels the filter left undecided (re-evaluated exactly) since the last reset."""
        n = C.c_ulonglong(0)
        self._check(self._lib.mod_filter_fallbacks(self._h, 1 if reset else 0, C.byref(n)))
        return int(n.value)

    def selfcheck_ddiv(self, a, b, stream: int = 0) -> int:
        """a, b: float64 CUDA tensors.  Returns the number of pairs where the kernel's division differs from IEEE."""
        import torch
        bad = torch.zeros(1, dtype=torch.int64, device=a.device)
        self._check(self._lib.mod_selfcheck_ddiv(self._h, stream or None, _dev_ptr(a), _dev_ptr(b), a.numel(),
                                                 _dev_ptr(bad)))
        return int(bad.item())

    def _check(self, rc: int, allow=()):
        if rc != _lib.MOD_OK and rc not in allow:
            raise ModError(rc, self._lib.mod_last_error(self._h).decode())
        return rc

    # ---- seam 1 ---------------------------------------------------------------------------------------------
    def scene_flow(self, now: DisparityImage, prev: DisparityImage, flow: np.ndarray, cam: CameraInfo,
                   prev2now: TransformMsg, dt_sec: float, config: SceneFlowConstructorConfig = None,
                   want_depth: bool = False, want_static_flow: bool = False, out: np.ndarray = None):
        """SceneFlowConstructor::construct (scene_flow_constructor.cpp:91-147) -> (points32[H,W,8], depth, static).
        `out`: a caller-owned (H,W,8) float32 array to fill instead of a fresh one (a message buffer kept across frames)."""
        config = config or SceneFlowConstructorConfig()
        h, w = now.image.shape
        assert flow.dtype == np.float32 and flow.shape == (h, w, 2) and flow.strides[1] == 8 and flow.strides[2] == 4
        if out is None:
            out = np.empty((h, w, 8), np.float32)
        assert out.dtype == np.float32 and out.shape == (h, w, 8) and out.flags.c_contiguous
        depth = np.empty((h, w), np.float32) if want_depth else None
        static = np.empty((h, w, 2), np.float32) if want_static_flow else None
        cn, cp, cc, ct, cfg = now.to_c(), prev.to_c(), cam.to_c(), prev2now.to_c(), config.to_c()
        self._check(self._lib.mod_scene_flow(
            self._h, w, h, C.byref(cc), C.byref(cn), C.byref(cp), flow.ctypes.data, flow.strides[0], C.byref(ct),
            float(dt_sec), C.byref(cfg), out.ctypes.data, depth.ctypes.data if want_depth else None,
            static.ctypes.data if want_static_flow else None))
        return out, depth, static

    # ---- seam 2 ---------------------------------------------------------------------------------------------
    def cluster(self, points32: np.ndarray, config: ClustererConfig = None, want_labels: bool = True,
                max_objs: int = 1024) -> ClusterResult:
        """ClustererNodelet::clustering + publishMovingObjects (clusterer_nodelet.cpp:85-95,324-343)."""
        config = config or ClustererConfig()
        assert points32.dtype == np.float32 and points32.ndim == 3 and points32.shape[2] == 8
        points32 = np.ascontiguousarray(points32)
        h, w = points32.shape[:2]
        labels = np.empty((h, w), np.int32) if want_labels else None
        objs = np.zeros(max_objs, OBJECT_DTYPE)
        n = C.c_int32(0)
        cfg = config.to_c()
        rc = self._check(self._lib.mod_cluster(self._h, w, h, points32.ctypes.data, C.byref(cfg),
                                               labels.ctypes.data if want_labels else None, objs.ctypes.data,
                                               max_objs, C.byref(n)), allow=(_lib.MOD_ERR_CAPACITY,))
        res = ClusterResult(objs[:min(n.value, max_objs)], labels)
        res.status = rc
        res.n_objects = n.value
        return res

    def cluster_resident(self, w: int, h: int, config: ClustererConfig = None, want_labels: bool = True,
                         max_objs: int = 1024) -> ClusterResult:
        """The clusterer on the cloud the last scene_flow() call left on the device (no PointCloud2 round trip)."""
        config = config or ClustererConfig()
        labels = np.empty((h, w), np.int32) if want_labels else None
        objs = np.zeros(max_objs, OBJECT_DTYPE)
        n = C.c_int32(0)
        cfg = config.to_c()
        rc = self._check(self._lib.mod_cluster_resident(self._h, w, h, C.byref(cfg),
                                                        labels.ctypes.data if want_labels else None, objs.ctypes.data,
                                                        max_objs, C.byref(n)), allow=(_lib.MOD_ERR_CAPACITY,))
        res = ClusterResult(objs[:min(n.value, max_objs)], labels)
        res.status = rc
        res.n_objects = n.value
        return res

    def clusters_image(self, labels: np.ndarray, n_clusters: int) -> np.ndarray:
        """ClustererNodelet::publishClustersImage (clusterer_nodelet.cpp:293-322): (H,W) int32 labels -> (H,W,3) BGR8."""
        labels = np.ascontiguousarray(labels, np.int32)
        h, w = labels.shape
        out = np.empty((h, w, 3), np.uint8)
        self._check(self._lib.mod_clusters_image(self._h, w, h, labels.ctypes.data, int(n_clusters), out.ctypes.data))
        return out

    def clusters_image_batched_dev(self, w: int, h: int, batch: int, labels, stats, out_bgr8, stream: int = 0):
        self._check(self._lib.mod_clusters_image_batched_dev(self._h, stream or None, w, h, batch, _dev_ptr(labels),
                                                             _dev_ptr(stats), _dev_ptr(out_bgr8)))

    # ---- image_crop (upstream glue) ---------------------------------------------------------------------------
    def image_crop(self, image: np.ndarray, encoding: str, target_w: int, target_h: int) -> np.ndarray:
        """ImageCrop::cameraCallback's image part (image_crop/src/image_crop.cpp:20-30): (H,W[,C]) uint8 in `encoding`
        (bgr8 / rgb8 / bgra8 / rgba8 / mono8; rows may be strided) -> (target_h, target_w, 3) bgr8."""
        assert image.dtype == np.uint8 and image.strides[-1] == 1
        h, w = image.shape[:2]
        row_bytes = image[0].size if h else 0
        step = image.strides[0] if h > 1 else row_bytes
        buf = np.zeros((h, step), np.uint8)  # sensor_msgs/Image.data: h rows of `step` bytes
        buf[:, :row_bytes] = image.reshape(h, row_bytes)
        out = np.empty((max(target_h, 0), max(target_w, 0), 3), np.uint8)
        self._check(self._lib.mod_image_crop(self._h, w, h, ENCODINGS[encoding], buf.ctypes.data, step, int(target_w),
                                             int(target_h), out.ctypes.data))
        return out

    def image_crop_batched_dev(self, w: int, h: int, batch: int, encoding: str, data, frame_stride_bytes: int,
                               step_bytes: int, target_w: int, target_h: int, out_bgr8, stream: int = 0):
        self._check(self._lib.mod_image_crop_batched_dev(self._h, stream or None, w, h, batch, ENCODINGS[encoding],
                                                         _dev_ptr(data), frame_stride_bytes, step_bytes, target_w,
                                                         target_h, _dev_ptr(out_bgr8)))

    # ---- MONO8 conversion (upstream glue) -----------------------------------------------------------------------
    def to_mono8(self, image: np.ndarray, encoding: str, rule: int = 0) -> np.ndarray:
        """cv_bridge::toCvCopy(image, MONO8) of estimateCameraMotion (scene_flow_constructor.cpp:220-221): (H,W[,C])
        uint8 in `encoding` (rows may be strided) -> (H,W) uint8.  rule: 0 = OpenCV 3.x grey constants (the
        reference's build), 1 = OpenCV 4.x."""
        assert image.dtype == np.uint8 and image.strides[-1] == 1
        h, w = image.shape[:2]
        row_bytes = image[0].size if h else 0
        step = image.strides[0] if h > 1 else row_bytes
        buf = np.zeros((h, step), np.uint8)
        buf[:, :row_bytes] = image.reshape(h, row_bytes)
        out = np.empty((h, w), np.uint8)
        self._check(self._lib.mod_image_to_mono8(self._h, w, h, ENCODINGS[encoding], buf.ctypes.data, step, int(rule),
                                                 out.ctypes.data))
        return out

    def to_mono8_batched_dev(self, w: int, h: int, batch: int, encoding: str, data, frame_stride_bytes: int,
                             step_bytes: int, rule: int, out_mono8, stream: int = 0):
        self._check(self._lib.mod_image_to_mono8_batched_dev(self._h, stream or None, w, h, batch, ENCODINGS[encoding],
                                                             _dev_ptr(data), frame_stride_bytes, step_bytes, int(rule),
                                                             _dev_ptr(out_mono8)))

    # ---- fused ----------------------------------------------------------------------------------------------
    def pipeline(self, now: DisparityImage, prev: DisparityImage, flow: np.ndarray, cam: CameraInfo,
                 prev2now: TransformMsg, dt_sec: float, sf_config: SceneFlowConstructorConfig = None,
                 cl_config: ClustererConfig = None, want_points: bool = False, want_labels: bool = True,
                 max_objs: int = 1024) -> ClusterResult:
        sf_config = sf_config or SceneFlowConstructorConfig()
        cl_config = cl_config or ClustererConfig()
        h, w = now.image.shape
        assert flow.dtype == np.float32 and flow.shape == (h, w, 2) and flow.strides[1] == 8
        pts = np.empty((h, w, 8), np.float32) if want_points else None
        labels = np.empty((h, w), np.int32) if want_labels else None
        objs = np.zeros(max_objs, OBJECT_DTYPE)
        n = C.c_int32(0)
        cn, cp, cc, ct = now.to_c(), prev.to_c(), cam.to_c(), prev2now.to_c()
        sfc, clc = sf_config.to_c(), cl_config.to_c()
        rc = self._check(self._lib.mod_pipeline(
            self._h, w, h, C.byref(cc), C.byref(cn), C.byref(cp), flow.ctypes.data, flow.strides[0], C.byref(ct),
            float(dt_sec), C.byref(sfc), C.byref(clc), pts.ctypes.data if want_points else None,
            labels.ctypes.data if want_labels else None, objs.ctypes.data, max_objs, C.byref(n)),
            allow=(_lib.MOD_ERR_CAPACITY,))
        res = ClusterResult(objs[:min(n.value, max_objs)], labels, pts)
        res.status = rc
        res.n_objects = n.value
        return res

    # ---- device-resident batched path (torch CUDA tensors; asynchronous on `stream`) ---------------------------
    @staticmethod
    def make_pairs(disp_now, disp_prev, flows, transforms: Sequence[TransformMsg], dts: Sequence[float], f: float,
                   T: float, min_disparity: float = 0.0, max_disparity: float = 128.0):
        """Builds the host array of mod_pair_dev for lists of dense device tensors (one entry per pair)."""
        n = len(transforms)
        arr = (PairDev * n)()
        for i in range(n):
            p = arr[i]
            p.disp_now, p.disp_prev, p.flow_xy = _dev_ptr(disp_now[i]), _dev_ptr(disp_prev[i]), _dev_ptr(flows[i])
            p.f_now = p.f_prev = f
            p.T_now = p.T_prev = T
            p.min_disparity_now = p.min_disparity_prev = min_disparity
            p.max_disparity_now = p.max_disparity_prev = max_disparity
            p.prev2now = transforms[i].to_c()
            p.dt_sec = float(dts[i])
        return arr

    def scene_flow_batched_dev(self, w: int, h: int, pairs, cam: CameraInfo, config: SceneFlowConstructorConfig,
                               dynamic_speed: float, points32, mask=None, depth=None, static_flow=None,
                               frame_stride_points: int = 0, stream: int = 0):
        cc, cfg = cam.to_c(), config.to_c()
        self._check(self._lib.mod_scene_flow_batched_dev(
            self._h, stream or None, w, h, len(pairs), C.byref(cc), C.cast(pairs, C.c_void_p), C.byref(cfg),
            float(dynamic_speed), _dev_ptr(points32), frame_stride_points or w * h, _dev_ptr(mask) or None,
            _dev_ptr(depth) or None, _dev_ptr(static_flow) or None))

    def cluster_batched_dev(self, w: int, h: int, batch: int, points32, mask, config: ClustererConfig, objs, stats,
                            max_objs: int, labels=None, frame_stride_points: int = 0, stream: int = 0):
        cfg = config.to_c()
        self._check(self._lib.mod_cluster_batched_dev(
            self._h, stream or None, w, h, batch, _dev_ptr(points32), frame_stride_points or w * h,
            _dev_ptr(mask) or None, C.byref(cfg), _dev_ptr(labels) or None, _dev_ptr(objs), max_objs,
            _dev_ptr(stats)))

    # ---- host streaming path -------------------------------------------------------------------------------------
    def pipeline_sequence_host(self, disp_frames, flows, cam: CameraInfo, transforms: Sequence[TransformMsg],
                               dts: Sequence[float], f: float, T: float, min_disparity: float, max_disparity: float,
                               sf_config: SceneFlowConstructorConfig, cl_config: ClustererConfig, objs, stats,
                               max_objs_per_pair: int, out_points32=None, transforms_c=None, dts_c=None):
        """disp_frames: (F,H,W) float32, flows: (F-1,H,W,2) float32, objs / stats: preallocated host arrays.
        Arrays may be numpy arrays or (pinned) torch CPU tensors."""
        def host_ptr(a):
            if a is None:
                return None
            return int(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data
        n_frames, h, w = disp_frames.shape
        n_pairs = n_frames - 1
        if transforms_c is None:
            transforms_c = (Transform * n_pairs)(*[t.to_c() for t in transforms])
            dts_c = (C.c_double * n_pairs)(*[float(d) for d in dts])
        cc, sfc, clc = cam.to_c(), sf_config.to_c(), cl_config.to_c()
        self._check(self._lib.mod_pipeline_sequence_host(
            self._h, w, h, n_frames, C.byref(cc), f, T, min_disparity, max_disparity, host_ptr(disp_frames),
            host_ptr(flows), C.cast(transforms_c, C.c_void_p), C.cast(dts_c, C.c_void_p), C.byref(sfc), C.byref(clc),
            host_ptr(out_points32), host_ptr(objs), max_objs_per_pair, host_ptr(stats)))
