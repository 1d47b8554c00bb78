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


def reproject(dmat: np.ndarray, gray, q: np.ndarray, depth_min: float, depth_max: float):
    d = np.ascontiguousarray(dmat, np.float32)
    H, W = d.shape
    depth = np.empty((H, W), np.float32)
    pts = np.zeros((H * W, 4), np.float32)
    g = np.ascontiguousarray(gray, np.uint8) if gray is not None else None
    qq = np.ascontiguousarray(q, np.float32)
    n = lib().sgbm_oracle_reproject(d.ctypes.data_as(ctypes.c_void_p),
                                    g.ctypes.data_as(ctypes.c_void_p) if g is not None else None,
                                    ctypes.c_int(W), ctypes.c_int(H), qq.ctypes.data_as(ctypes.c_void_p),
                                    ctypes.c_float(depth_min), ctypes.c_float(depth_max),
                                    depth.ctypes.data_as(ctypes.c_void_p), pts.ctypes.data_as(ctypes.c_void_p))
    return depth, pts[:n].copy()
