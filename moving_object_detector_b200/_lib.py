"""ctypes binding of the C ABI declared in include/mod_b200.h (libmod_b200.so, built by csrc/Makefile).

The library is the product: there is no Python / CPU fallback.  `load()` raises if the shared object is missing,
and every compute entry point returns MOD_ERR_NO_DEVICE / MOD_ERR_CUDA (surfaced as `ModError`) without a B200.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOD_B200_LIB lets a developer point the harness at an experimental build of the same C ABI (kernel tuning sweeps)
LIB_PATH = os.environ.get("MOD_B200_LIB") or os.path.join(_HERE, "libmod_b200.so")

MOD_OK = 0
MOD_ERR_INVALID_ARG = -1
MOD_ERR_CAPACITY = -2
MOD_ERR_INVALID_PARAM = -3
MOD_ERR_CUDA = -4
MOD_ERR_NO_DEVICE = -5
MOD_ERR_CLUSTER_SLOTS = -6

_STATUS_NAMES = {
    MOD_OK: "MOD_OK",
    MOD_ERR_INVALID_ARG: "MOD_ERR_INVALID_ARG",
    MOD_ERR_CAPACITY: "MOD_ERR_CAPACITY",
    MOD_ERR_INVALID_PARAM: "MOD_ERR_INVALID_PARAM",
    MOD_ERR_CUDA: "MOD_ERR_CUDA",
    MOD_ERR_NO_DEVICE: "MOD_ERR_NO_DEVICE",
    MOD_ERR_CLUSTER_SLOTS: "MOD_ERR_CLUSTER_SLOTS",
}


class ModError(RuntimeError):
    def __init__(self, status: int, message: str = ""):
        self.status = status
        name = _STATUS_NAMES.get(status, str(status))
        super().__init__(f"{name}: {message}" if message else name)


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fx", "fy", "cx", "cy", "Tx", "Ty")]


class Disparity(C.Structure):
    _fields_ = [("f", C.c_float), ("T", C.c_float), ("min_disparity", C.c_float), ("max_disparity", C.c_float),
                ("data", C.c_void_p), ("step_bytes", C.c_int32)]


class Transform(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("tx", "ty", "tz", "qx", "qy", "qz", "qw")]


class SfParams(C.Structure):
    _fields_ = [("dynamic_flow_diff", C.c_int32), ("max_color_velocity", C.c_double)]


class ClParams(C.Structure):
    _fields_ = [("cluster_size", C.c_int32), ("depth_diff", C.c_double), ("dynamic_speed", C.c_double),
                ("neighbor_distance", C.c_int32)]


class Object(C.Structure):
    _fields_ = [("id", C.c_int32), ("center", C.c_float * 3), ("velocity", C.c_float * 3), ("bbox", C.c_float * 3),
                ("n_points", C.c_int32)]


class PairDev(C.Structure):
    _fields_ = [("disp_now", C.c_void_p), ("disp_prev", C.c_void_p), ("flow_xy", C.c_void_p),
                ("f_now", C.c_float), ("T_now", C.c_float), ("min_disparity_now", C.c_float),
                ("max_disparity_now", C.c_float),
                ("f_prev", C.c_float), ("T_prev", C.c_float), ("min_disparity_prev", C.c_float),
                ("max_disparity_prev", C.c_float),
                ("prev2now", Transform), ("dt_sec", C.c_double)]


class FrameStats(C.Structure):
    _fields_ = [("n_dynamic", C.c_int32), ("n_clusters", C.c_int32), ("n_objects", C.c_int32), ("status", C.c_int32)]


# Every symbol include/mod_b200.h declares, with its prototype.  tests/test_abi_symbols.py checks this list against
# the header and against the built library.
PROTOTYPES = {
    "mod_abi_version": (C.c_int, []),
    "mod_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "mod_destroy": (None, [C.c_void_p]),
    "mod_last_error": (C.c_char_p, [C.c_void_p]),
    "mod_launch_count": (C.c_ulonglong, [C.c_void_p]),
    "mod_stream": (C.c_void_p, [C.c_void_p]),
    "mod_last_frame_stats": (C.c_int, [C.c_void_p, C.POINTER(FrameStats)]),
    "mod_max_clusters": (C.c_int, [C.c_void_p]),
    "mod_set_static_flow_filter": (C.c_int, [C.c_void_p, C.c_int]),
    "mod_filter_fallbacks": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]),
    "mod_selfcheck_ddiv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mod_host_rotation": (None, [C.POINTER(Transform), C.c_void_p, C.c_void_p]),
    "mod_host_threshold_ge": (C.c_float, [C.c_double]),
    "mod_host_threshold_gt": (C.c_float, [C.c_double]),
    "mod_host_sq_threshold": (C.c_float, [C.c_float]),
    "mod_host_alloc": (C.c_void_p, [C.c_size_t]),
    "mod_host_free": (None, [C.c_void_p]),
    "mod_host_register": (C.c_int, [C.c_void_p, C.c_size_t]),
    "mod_host_unregister": (C.c_int, [C.c_void_p]),
    "mod_set_host_copy_threads": (None, [C.c_int]),
    "mod_host_copy_threads": (C.c_int, []),
    "mod_set_frame_ids": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_ulonglong]),
    "mod_resident_hits": (C.c_ulonglong, [C.c_void_p]),
    "mod_scene_flow": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(Camera), C.POINTER(Disparity),
                                 C.POINTER(Disparity), C.c_void_p, C.c_int32, C.POINTER(Transform), C.c_double,
                                 C.POINTER(SfParams), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mod_cluster": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(ClParams), C.c_void_p,
                              C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "mod_cluster_resident": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(ClParams), C.c_void_p, C.c_void_p,
                                       C.c_int32, C.POINTER(C.c_int32)]),
    "mod_pipeline": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(Camera), C.POINTER(Disparity),
                               C.POINTER(Disparity), C.c_void_p, C.c_int32, C.POINTER(Transform), C.c_double,
                               C.POINTER(SfParams), C.POINTER(ClParams), C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int32, C.POINTER(C.c_int32)]),
    "mod_scene_flow_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Camera),
                                             C.c_void_p, C.POINTER(SfParams), C.c_double, C.c_void_p, C.c_size_t,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "mod_cluster_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                          C.c_void_p, C.POINTER(ClParams), C.c_void_p, C.c_void_p, C.c_int32,
                                          C.c_void_p]),
    "mod_clusters_image": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int32, C.c_void_p]),
    "mod_clusters_image_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
    "mod_host_hsv_colormap": (None, [C.c_void_p]),
    "mod_image_crop": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int32, C.c_int, C.c_int,
                                 C.c_void_p]),
    "mod_image_crop_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_size_t, C.c_int32, C.c_int, C.c_int, C.c_void_p]),
    "mod_host_encoding": (C.c_int, [C.c_char_p]),
    "mod_image_to_mono8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int32, C.c_int, C.c_void_p]),
    "mod_image_to_mono8_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                 C.c_size_t, C.c_int32, C.c_int, C.c_void_p]),
    "mod_host_crop_camera_info": (C.c_int, [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "mod_pipeline_sequence_host": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Camera), C.c_float,
                                             C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.POINTER(SfParams), C.POINTER(ClParams), C.c_void_p,
                                             C.c_void_p, C.c_int32, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Loads libmod_b200.so (once) and attaches prototypes.  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C moving_object_detector_b200/csrc` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib
