"""TEST INFRASTRUCTURE: ctypes wrapper over oracle/sgbm_oracle.c (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libsgbm_oracle.so")
_lib = None

_FIELDS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff",
           "preFilterCap", "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")


class _CParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in _FIELDS]


class _CDumps(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("C", "S", "disp_wta", "disp_med")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "sgbm_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.sgbm_oracle_compute.restype = ctypes.c_int
        _lib.sgbm_oracle_reproject.restype = ctypes.c_uint32
    return _lib


def _cparams(p) -> _CParams:
    return _CParams(**{n: int(getattr(p, n)) for n in _FIELDS})


def compute(left: np.ndarray, right: np.ndarray, p, dumps: bool = False):
    """Returns disp (H,W) int16; with dumps=True also a dict of stage outputs."""
    L = np.ascontiguousarray(left, np.uint8)
    R = np.ascontiguousarray(right, np.uint8)
    H, W = L.shape
    disp = np.empty((H, W), np.int16)
    cp = _cparams(p)
    dd = None
    out = {}
    if dumps:
        D = p.numDisparities
        W1 = max(p.w1(W), 0)
        out = dict(C=np.zeros((H, W1, D), np.int16), S=np.zeros((H, W1, D), np.int16),
                   disp_wta=np.empty((H, W), np.int16), disp_med=np.empty((H, W), np.int16))
        dd = _CDumps(*[out[k].ctypes.data for k in ("C", "S", "disp_wta", "disp_med")])
    rc = lib().sgbm_oracle_compute(L.ctypes.data_as(ctypes.c_void_p), R.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.c_int(W), ctypes.c_int(H), ctypes.byref(cp),
                                   disp.ctypes.data_as(ctypes.c_void_p),
                                   ctypes.byref(dd) if dd is not None else None)
    if rc != 0:
        raise ValueError("sgbm_oracle_compute failed rc=%d" % rc)
    return (disp, out) if dumps else disp


def median3x3(img: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(img, np.int16)
    o = np.empty_like(a)
    lib().sgbm_oracle_median3x3(a.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_int(a.shape[1]), ctypes.c_int(a.shape[0]))
    return o


def filter_speckles(img: np.ndarray, new_val: int, max_size: int, max_diff: int) -> np.ndarray:
    a = np.array(img, np.int16, order="C", copy=True)
    lib().sgbm_oracle_filter_speckles(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(a.shape[1]), ctypes.c_int(a.shape[0]),
                                      ctypes.c_int(new_val), ctypes.c_int(max_size), ctypes.c_int(max_diff))
    return a


def to_float(disp16: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(disp16, np.int16)
    o = np.empty(a.shape, np.float32)
    lib().sgbm_oracle_to_float(a.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.size))
    return o


def process_disparity(disp16: np.ndarray, min_disp: float, max_disp: float) -> np.ndarray:
    a = np.ascontiguousarray(disp16, np.int16)
    o = np.empty(a.shape, np.float32)
    lib().sgbm_oracle_process_disparity(a.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.c_size_t(a.size), ctypes.c_float(min_disp), ctypes.c_float(max_disp))
    return o


def calc_q(fx: float, cx: float, cxr: float, cy: float, p14: float) -> np.ndarray:
    """calc_q of disparity_to_depth.cpp:62-85 in double, cast to float32 as :136-140 does.
    Returns [q03, q13, wz, q32, q33]."""
    T = -p14 / fx
    q33 = -(cx - cxr) / T
    return np.array([-cx, -cy, fx, 1.0 / T, q33], np.float64).astype(np.float32)


def reproject(dmat: np.ndarray, color, q: np.ndarray, depth_min: float, depth_max: float):
    """Row R.  `color`: (H, W) MONO8 or (H, W, 3) BGR8 (disparity_to_depth.cpp:111-125, :176-188), or None."""
    d = np.ascontiguousarray(dmat, np.float32)
    H, W = d.shape
    depth = np.empty((H, W), np.float32)
    pts = np.zeros((H * W, 4), np.float32)
    g = np.ascontiguousarray(color, np.uint8) if color is not None else None
    ch = 0 if g is None else (1 if g.ndim == 2 else 3)
    qq = np.ascontiguousarray(q, np.float32)
    n = lib().sgbm_oracle_reproject(d.ctypes.data_as(ctypes.c_void_p),
                                    g.ctypes.data_as(ctypes.c_void_p) if g is not None else None, ctypes.c_int(ch),
                                    ctypes.c_int(W), ctypes.c_int(H), qq.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_double(depth_min), ctypes.c_double(depth_max),
                                    depth.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p))
    return depth, pts[:n].copy()


def disparity_window(fx: float, baseline: float, depth_min: float, depth_max: float):
    """min_disparity / max_disparity exactly as generate_disparity.cpp:441-450 forms them: f and T are float32 message fields,
    their product is a float32, the division by the double depth bound happens in double and the result is stored as float32
    (depth_min == 0 gives +inf)."""
    f, T = np.float32(fx), np.float32(baseline)
    tf = np.float32(T * f)
    with np.errstate(divide="ignore"):
        lo = np.float32(np.float64(tf) / np.float64(depth_max))
        hi = np.float32(np.float64(tf) / np.float64(depth_min)) if depth_min != 0 else np.float32(np.inf)
    return float(lo), float(hi)


# ---- oracle/_ref: the reference's own reprojection statements compiled against stand-in types (oracle/build_ref.py) ----
_ref = None


def ref_lib():
    """libref_reproject.so, or None when it can be neither built (no /root/reference) nor found prebuilt."""
    global _ref
    if _ref is None:
        from . import build_ref
        path = build_ref.build()
        if path is None:
            return None
        _ref = ctypes.CDLL(path)
        _ref.ref_reproject.restype = ctypes.c_uint
    return _ref


def ref_calc_q(Kl, Pr, Pl) -> np.ndarray:
    k, pr, pl = (np.ascontiguousarray(a, np.float64) for a in (Kl, Pr, Pl))
    q = np.zeros(16, np.float64)
    ref_lib().ref_calc_q(k.ctypes.data_as(ctypes.c_void_p), pr.ctypes.data_as(ctypes.c_void_p), pl.ctypes.data_as(ctypes.c_void_p),
                         q.ctypes.data_as(ctypes.c_void_p))
    return q.reshape(4, 4)


def ref_process_disparity(disp32: np.ndarray, fx: float, baseline: float, depth_min: float, depth_max: float):
    a = np.ascontiguousarray(disp32, np.float32)
    out = np.empty_like(a)
    mm = np.zeros(2, np.float32)
    ref_lib().ref_process_disparity(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(a.shape[0]), ctypes.c_int(a.shape[1]),
                                    ctypes.c_double(fx), ctypes.c_double(baseline), ctypes.c_double(depth_min), ctypes.c_double(depth_max),
                                    out.ctypes.data_as(ctypes.c_void_p), mm.ctypes.data_as(ctypes.c_void_p))
    return out, float(mm[0]), float(mm[1])


def ref_reproject(dmat: np.ndarray, color: np.ndarray, Kl, Pr, Pl, depth_min: float, depth_max: float):
    d = np.ascontiguousarray(dmat, np.float32)
    H, W = d.shape
    g = np.ascontiguousarray(color, np.uint8)
    ch = 1 if g.ndim == 2 else 3
    k, pr, pl = (np.ascontiguousarray(a, np.float64) for a in (Kl, Pr, Pl))
    depth = np.empty((H, W), np.float32)
    pts = np.zeros((H * W, 4), np.float32)
    n = ref_lib().ref_reproject(d.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(H), ctypes.c_int(W), g.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_int(ch), k.ctypes.data_as(ctypes.c_void_p), pr.ctypes.data_as(ctypes.c_void_p),
                                pl.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(depth_min), ctypes.c_double(depth_max),
                                ctypes.c_float(0), ctypes.c_float(0), depth.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p))
    return depth, pts[:n].copy()
