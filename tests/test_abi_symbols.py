"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/mod_b200.h declares, and fails
loudly (no CPU fallback) when asked to compute without a B200.  No compute call is made here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mod_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mod_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("mod_create", "mod_destroy", "mod_scene_flow", "mod_cluster", "mod_pipeline",
                 "mod_scene_flow_batched_dev", "mod_cluster_batched_dev", "mod_pipeline_sequence_host"):
        assert must in names


def test_library_exports_every_declared_symbol(mod):
    from moving_object_detector_b200 import _lib
    lib = mod.load()
    names = declared_symbols()
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and header are out of sync"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mod_b200.h but not exported by libmod_b200.so"
    out = subprocess.check_output(["nm", "-D", "--defined-only", mod.LIB_PATH], text=True)
    exported = set(re.findall(r" T (mod_[a-z0-9_]+)", out))
    assert set(names) <= exported


def test_abi_version(mod):
    assert mod.load().mod_abi_version() == 1


def test_struct_layouts_match_the_header(mod):
    from moving_object_detector_b200 import _lib
    assert C.sizeof(_lib.Camera) == 48 and C.sizeof(_lib.Transform) == 56
    assert C.sizeof(_lib.Object) == 44 and C.sizeof(_lib.FrameStats) == 16
    assert C.sizeof(_lib.Disparity) == 32 and C.sizeof(_lib.ClParams) == 32 and C.sizeof(_lib.SfParams) == 16
    assert C.sizeof(_lib.PairDev) == 24 + 32 + 56 + 8
    assert mod.OBJECT_DTYPE.itemsize == 44 and mod.STATS_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_a_gpu(mod):
    """mod_create must refuse (MOD_ERR_NO_DEVICE) on a box without a CUDA device; the Python harness raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from moving_object_detector_b200 import _lib
    h = C.c_void_p()
    rc = mod.load().mod_create(0, 64, 64, 1, 0, C.byref(h))
    assert rc == _lib.MOD_ERR_NO_DEVICE and not h.value
    with pytest.raises(mod.ModError):
        mod.Context(64, 64)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package (Python or C++) may reference it."""
    pkg = os.path.join(ROOT, "moving_object_detector_b200")
    for base, _, files in os.walk(pkg):
        if "build" in base.split(os.sep):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or fn == "Makefile":
                text = open(os.path.join(base, fn), errors="ignore").read()
                for pattern in (r"import\s+oracle_py|from\s+oracle", r"#include\s*[<\"][^>\"]*mod_oracle",
                                r"libmod_oracle", r"\borc_[a-z_]+\s*\("):
                    assert not re.search(pattern, text), f"{fn} uses the oracle ({pattern})"
    out = subprocess.check_output(["nm", "-D", os.path.join(pkg, "libmod_b200.so")], text=True)
    assert "orc_" not in out


# ---- O(1) host-side parameter preparation (runs on the CPU, part of the product) ----------------------------------
def test_host_rotation_matches_eigen_order(mod, oracle):
    import np_restatement as npr
    from moving_object_detector_b200 import _lib
    rng = np.random.RandomState(3)
    for _ in range(50):
        q = rng.normal(0, 1, 4) * rng.uniform(0.5, 2.0)       # not normalised on purpose
        t = rng.normal(0, 1, 3)
        tr = _lib.Transform(t[0], t[1], t[2], q[0], q[1], q[2], q[3])
        R, tt = np.empty(9), np.empty(3)
        mod.load().mod_host_rotation(C.byref(tr), R.ctypes.data, tt.ctypes.data)
        assert R.reshape(3, 3).tobytes() == npr.rotation_from_quaternion(q).tobytes()
        assert R.reshape(3, 3).tobytes() == oracle.quat_to_matrix(np.concatenate([t, q])).tobytes()
        assert tt.tobytes() == t.tobytes()


def test_float_threshold_equivalents(mod):
    lib = mod.load()
    rng = np.random.RandomState(4)
    ths = list(rng.uniform(0, 10, 200)) + [0.3, 0.15, 0.0, 1.0, 0.1, 2.5, 1e-40, 3.0000001]
    for th in ths:
        ge, gt = np.float32(lib.mod_host_threshold_ge(th)), np.float32(lib.mod_host_threshold_gt(th))
        for a in (np.nextafter(ge, np.float32(-np.inf)), ge, np.nextafter(ge, np.float32(np.inf))):
            assert (np.float64(a) >= th) == (a >= ge)
        for a in (np.nextafter(gt, np.float32(-np.inf)), gt, np.nextafter(gt, np.float32(np.inf))):
            assert (np.float64(a) > th) == (a > gt)


def test_sqrt_free_threshold(mod):
    lib = mod.load()
    for thr in [0.0, -1.0, 1.0, 2.0, 3.0, 5.0, 7.0, 100.0, 0.3, 0.30000001192092896, 1e-20, 1e19, 1.1754944e-38]:
        thr = np.float32(thr)
        s2 = np.float32(lib.mod_host_sq_threshold(thr))
        cand = [s2]
        lo, hi = s2, s2
        for _ in range(4):
            lo = np.nextafter(lo, np.float32(0.0))
            hi = np.nextafter(hi, np.float32(np.inf))
            cand += [lo, hi]
        for s in cand + [np.float32(0.0), np.float32(np.inf)]:
            assert (np.sqrt(s) >= thr) == (s >= s2), (thr, s, s2)
    assert np.isnan(lib.mod_host_sq_threshold(np.float32(np.nan)))


def test_host_copy_thread_setting_needs_no_gpu(mod):
    """mod_set_host_copy_threads / mod_host_copy_threads are host-only (csrc/hostcopy.cu): the pool can be sized,
    clamped to [0, 16] and switched off without a device; frame ids and hits are per-context and need one."""
    lib = mod.load()
    before = lib.mod_host_copy_threads()
    try:
        for n, want in ((3, 3), (1, 1), (0, 0), (-5, 0), (100, 16), (4, 4)):
            lib.mod_set_host_copy_threads(n)
            assert lib.mod_host_copy_threads() == want
    finally:
        lib.mod_set_host_copy_threads(before)
    assert lib.mod_set_frame_ids(None, 1, 2) != 0 and lib.mod_resident_hits(None) == 0
    assert lib.mod_host_register(None, 0) != 0
