"""ctypes binding of libb200sgm.so (the C ABI of include/b200sgm.h).

There is no CPU fallback: importing works without a GPU, but creating an Engine raises if the CUDA
library is missing or no device is usable.
"""
from __future__ import annotations

import ctypes
import os
import numpy as np

from .params import CParams, SGBMParams

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200SGM_LIB") or os.path.join(_HERE, "libb200sgm.so")   # B200SGM_LIB: development builds (build.py variants)
_lib = None

# every symbol include/b200sgm.h declares
SYMBOLS = (
    "b200sgm_create", "b200sgm_destroy", "b200sgm_set_params", "b200sgm_get_effective_params", "b200sgm_compute",
    "b200sgm_compute_f32", "b200sgm_compute_device", "b200sgm_enqueue", "b200sgm_wait", "b200sgm_compute_xyz",
    "b200sgm_last_error", "b200sgm_version", "b200sgm_launch_count", "b200sgm_lane_stream", "b200sgm_debug_read",
    "b200sgm_debug_set_path", "b200sgm_debug_overlap", "b200sgm_profile", "b200sgm_stage_times", "b200sgm_alu_peak", "b200sgm_stage_timeline",
    "b200sgm_set_camera", "b200sgm_rectify", "b200sgm_rectify_device", "b200sgm_rectify_maps", "b200sgm_bm_compute",
    "b200sgm_bm_compute_device", "b200sgm_lane_status", "b200sgm_reproject_from_camera",
    "b200sgm_create_bm", "b200sgm_host_alloc", "b200sgm_host_free", "b200sgm_bm_compute_f32",
)

STAGES = ("prefilter", "cost", "horizontal", "vertical_wta", "lrcheck", "median", "speckle")


class B200SGMError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200sgm error %d: %s" % (code, msg))
        self.code = code


class CReproject(ctypes.Structure):
    """Mirror of `b200sgm_reproject` in include/b200sgm.h."""
    _fields_ = [(n, ctypes.c_float) for n in ("q03", "q13", "wz", "q32", "q33", "min_disparity", "max_disparity")] + \
               [("depth_min", ctypes.c_double), ("depth_max", ctypes.c_double), ("color", ctypes.c_void_p),
                ("color_stride", ctypes.c_size_t), ("color_channels", ctypes.c_int)]


def load_library():
    """Loads the in-tree CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libb200sgm.so is not built: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                           "There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for s in SYMBOLS:
        getattr(lib, s)
    lib.b200sgm_last_error.restype = ctypes.c_char_p
    lib.b200sgm_version.restype = ctypes.c_char_p
    _lib = lib
    return lib


def _u8(a):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("images must be 2-D uint8 (CV_8UC1)")
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


class Engine:
    """One engine per GPU. `lanes` = frames that may be in flight (each lane owns its scratch volumes)."""

    def __init__(self, device=0, max_width=640, max_height=480, max_disparities=64, lanes=1, params: SGBMParams | None = None,
                 bm_only=False):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        create = self.lib.b200sgm_create_bm if bm_only else self.lib.b200sgm_create
        rc = create(int(device), int(max_width), int(max_height), int(max_disparities), int(lanes), ctypes.byref(self.h))
        if rc != 0:
            raise B200SGMError(rc, "b200sgm_create failed (no usable CUDA device or out of memory)")
        self.device = device
        self.lanes = lanes
        self.params = None
        self.last_warning = None
        if params is not None:
            self.set_params(params)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200sgm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        """Negative codes raise; positive codes are warnings (the result was delivered): kept in `last_warning`."""
        if rc < 0:
            raise B200SGMError(rc, self.lib.b200sgm_last_error(self.h).decode())
        self.last_warning = (rc, self.lib.b200sgm_last_error(self.h).decode()) if rc > 0 else None

    def lane_status(self, lane=0) -> int:
        """Status of the last frame of `lane` (after the caller synchronised its stream): 0, WARN_COST_RANGE (1); raises on error."""
        rc = self.lib.b200sgm_lane_status(self.h, int(lane))
        self._check(rc)
        return rc

    def set_params(self, p: SGBMParams):
        cp = p.to_c()
        self._check(self.lib.b200sgm_set_params(self.h, ctypes.byref(cp)))
        self.params = p

    def effective_params(self) -> SGBMParams:
        cp = CParams()
        self._check(self.lib.b200sgm_get_effective_params(self.h, ctypes.byref(cp)))
        return SGBMParams(**{n: getattr(cp, n) for n, _ in CParams._fields_})

    # ---- host-pointer synchronous path (what the matcher plugin calls) ----
    def compute(self, left, right) -> np.ndarray:
        L, R = _u8(left), _u8(right)
        if L.shape != R.shape:
            raise ValueError("Images MUST be the same resolution")
        H, W = L.shape
        disp = np.empty((H, W), np.int16)
        self._check(self.lib.b200sgm_compute(self.h, L.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(L.strides[0]),
                                             R.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(R.strides[0]),
                                             W, H, disp.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(W * 2)))
        return disp

    def compute_f32(self, left, right) -> np.ndarray:
        L, R = _u8(left), _u8(right)
        H, W = L.shape
        out = np.empty((H, W), np.float32)
        self._check(self.lib.b200sgm_compute_f32(self.h, L.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(L.strides[0]),
                                                 R.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(R.strides[0]),
                                                 W, H, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(W * 4)))
        return out

    def compute_xyz(self, left, right, q, depth_min, depth_max, min_disparity, max_disparity, want_points=True, color=None, out=None):
        """Matching + processDisparity + reprojection.  `color`: None (the left image, MONO8), (H, W) uint8 or (H, W, 3) BGR8."""
        L, R = _u8(left), _u8(right)
        H, W = L.shape
        rp = CReproject(float(q[0]), float(q[1]), float(q[2]), float(q[3]), float(q[4]), float(min_disparity), float(max_disparity),
                        float(depth_min), float(depth_max), None, 0, 0)
        if color is not None:
            color = np.ascontiguousarray(color, np.uint8)
            if color.shape[:2] != (H, W) or color.ndim not in (2, 3) or (color.ndim == 3 and color.shape[2] != 3):
                raise ValueError("color must be (H, W) MONO8 or (H, W, 3) BGR8 of the frame's size")
            rp.color = color.ctypes.data
            rp.color_stride = color.strides[0]
            rp.color_channels = 1 if color.ndim == 2 else 3
        if out is None:
            disp = np.empty((H, W), np.int16)
            dmat = np.empty((H, W), np.float32)
            depth = np.empty((H, W), np.float32)
            pts = np.empty((H * W, 4), np.float32) if want_points else None
        else:
            disp, dmat, depth, pts = out      # caller-owned (e.g. page-locked) arrays of exactly these shapes and dtypes
            assert disp.dtype == np.int16 and dmat.dtype == np.float32 and depth.dtype == np.float32 and disp.shape == (H, W)
            assert pts is None or (pts.dtype == np.float32 and pts.shape == (H * W, 4))
        cnt = ctypes.c_uint32(0)
        self._check(self.lib.b200sgm_compute_xyz(
            self.h, L.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(L.strides[0]), R.ctypes.data_as(ctypes.c_void_p),
            ctypes.c_size_t(R.strides[0]), W, H, ctypes.byref(rp), disp.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(W * 2),
            dmat.ctypes.data_as(ctypes.c_void_p), depth.ctypes.data_as(ctypes.c_void_p),
            pts.ctypes.data_as(ctypes.c_void_p) if pts is not None else None, ctypes.byref(cnt)))
        return disp, dmat, depth, (pts[:cnt.value] if pts is not None else None), cnt.value

    # ---- rectification (row N2): cv::initUndistortRectifyMap + cv::remap(INTER_CUBIC, BORDER_CONSTANT) ----
    def set_camera(self, cam, K, D, R, P):
        K = np.ascontiguousarray(K, np.float64).reshape(9)
        D = np.ascontiguousarray(D if D is not None else [], np.float64).ravel()
        P = np.ascontiguousarray(P, np.float64).reshape(12)
        Rp = None
        if R is not None:
            R = np.ascontiguousarray(R, np.float64).reshape(9)
            Rp = R.ctypes.data_as(ctypes.c_void_p)
        self._check(self.lib.b200sgm_set_camera(self.h, int(cam), K.ctypes.data_as(ctypes.c_void_p),
                                                D.ctypes.data_as(ctypes.c_void_p) if D.size else None, int(D.size), Rp,
                                                P.ctypes.data_as(ctypes.c_void_p)))

    def rectify(self, cam, image) -> np.ndarray:
        a = _u8(image)
        H, W = a.shape
        out = np.empty((H, W), np.uint8)
        self._check(self.lib.b200sgm_rectify(self.h, int(cam), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(a.strides[0]), W, H,
                                             out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(W)))
        return out

    def rectify_device(self, lane, cam, sptr, sstride, W, H, dptr, dstride, stream=0):
        self._check(self.lib.b200sgm_rectify_device(self.h, int(lane), int(cam), ctypes.c_void_p(sptr), ctypes.c_size_t(sstride), int(W),
                                                    int(H), ctypes.c_void_p(dptr), ctypes.c_size_t(dstride), ctypes.c_void_p(stream)))

    def rectify_maps(self, cam, W, H):
        m1 = np.empty((H, W), np.float32)
        m2 = np.empty((H, W), np.float32)
        self._check(self.lib.b200sgm_rectify_maps(self.h, int(cam), int(W), int(H), m1.ctypes.data_as(ctypes.c_void_p),
                                                  m2.ctypes.data_as(ctypes.c_void_p)))
        return m1, m2

    # ---- StereoBM (row N4): cv::StereoBM::compute of matcherOpenCVBlock.cpp:13-20 ----
    def bm_compute(self, left, right, numDisparities=64, blockSize=9, minDisparity=0, preFilterCap=31, textureThreshold=10,
                   uniquenessRatio=15, speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1, preFilterSize=0, f32=False) -> np.ndarray:
        L, R = _u8(left), _u8(right)
        if L.shape != R.shape:
            raise ValueError("Images MUST be the same resolution")
        H, W = L.shape
        bp = (ctypes.c_int * 10)(minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio,
                                 speckleWindowSize, speckleRange, disp12MaxDiff, preFilterSize)
        out = np.empty((H, W), np.float32 if f32 else np.int16)
        fn = self.lib.b200sgm_bm_compute_f32 if f32 else self.lib.b200sgm_bm_compute
        self._check(fn(self.h, bp, L.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(L.strides[0]),
                       R.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(R.strides[0]), W, H,
                       out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(out.strides[0])))
        return out

    def bm_compute_device(self, lane, lptr, lstride, rptr, rstride, W, H, dptr, dstride, stream=0, numDisparities=64, blockSize=9,
                          minDisparity=0, preFilterCap=31, textureThreshold=10, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0):
        bp = (ctypes.c_int * 10)(minDisparity, numDisparities, blockSize, preFilterCap, textureThreshold, uniquenessRatio,
                                 speckleWindowSize, speckleRange, -1, 0)
        self._check(self.lib.b200sgm_bm_compute_device(self.h, int(lane), bp, ctypes.c_void_p(lptr), ctypes.c_size_t(lstride),
                                                       ctypes.c_void_p(rptr), ctypes.c_size_t(rstride), int(W), int(H),
                                                       ctypes.c_void_p(dptr), ctypes.c_size_t(dstride), ctypes.c_void_p(stream)))

    # ---- streaming with host buffers (pinned for true overlap) ----
    def enqueue(self, lane, left, right, disp_out):
        L, R = _u8(left), _u8(right)
        H, W = L.shape
        assert disp_out.dtype == np.int16 and disp_out.shape == (H, W) and disp_out.strides[1] == 2
        self._check(self.lib.b200sgm_enqueue(self.h, int(lane), L.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(L.strides[0]),
                                             R.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(R.strides[0]), W, H,
                                             disp_out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(disp_out.strides[0])))

    def enqueue_ptr(self, lane, lptr, lstride, rptr, rstride, W, H, dptr, dstride):
        self._check(self.lib.b200sgm_enqueue(self.h, int(lane), ctypes.c_void_p(lptr), ctypes.c_size_t(lstride), ctypes.c_void_p(rptr),
                                             ctypes.c_size_t(rstride), int(W), int(H), ctypes.c_void_p(dptr), ctypes.c_size_t(dstride)))

    def wait(self, lane):
        self._check(self.lib.b200sgm_wait(self.h, int(lane)))

    # ---- device-resident path (raw device pointers, e.g. torch tensors' data_ptr()) ----
    def compute_device(self, lane, lptr, lstride, rptr, rstride, W, H, dptr, dstride, stream=0):
        self._check(self.lib.b200sgm_compute_device(self.h, int(lane), ctypes.c_void_p(lptr), ctypes.c_size_t(lstride),
                                                    ctypes.c_void_p(rptr), ctypes.c_size_t(rstride), int(W), int(H),
                                                    ctypes.c_void_p(dptr), ctypes.c_size_t(dstride), ctypes.c_void_p(stream)))

    def lane_stream(self, lane) -> int:
        s = ctypes.c_void_p()
        self._check(self.lib.b200sgm_lane_stream(self.h, int(lane), ctypes.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        c = ctypes.c_uint64(0)
        self._check(self.lib.b200sgm_launch_count(self.h, ctypes.byref(c)))
        return c.value

    def profile(self, enable: bool):
        self._check(self.lib.b200sgm_profile(self.h, int(bool(enable))))

    def stage_times(self, lane=0):
        """(dict stage -> accumulated ms, frames) since the previous call; synchronises the device."""
        ms = (ctypes.c_double * len(STAGES))()
        fr = ctypes.c_uint64(0)
        self._check(self.lib.b200sgm_stage_times(self.h, int(lane), ms, len(STAGES), ctypes.byref(fr)))
        return {k: ms[i] for i, k in enumerate(STAGES)}, fr.value

    def stage_timeline(self, lane=0):
        """(frames, 7) array of stage-boundary timestamps in ms since profile(True); call stage_times first."""
        n = ctypes.c_int(0)
        self._check(self.lib.b200sgm_stage_timeline(self.h, int(lane), None, 0, ctypes.byref(n)))
        buf = np.zeros(n.value, np.float32)
        self._check(self.lib.b200sgm_stage_timeline(self.h, int(lane), buf.ctypes.data_as(ctypes.c_void_p), n.value, ctypes.byref(n)))
        return buf.reshape(-1, len(STAGES) + 1)

    def set_path(self, path: int):
        self._check(self.lib.b200sgm_debug_set_path(self.h, int(path)))

    def debug_volume(self, what, W, H, lane=0) -> np.ndarray:
        """Reads the C or S cost volume of the last frame as [H][W1][D] uint16 (padding stripped)."""
        p = self.params
        W1 = p.w1(W)
        n = 1
        while 64 * n < p.numDisparities:
            n *= 2
        Dp = (p.numDisparities + 2 * n - 1) // (2 * n) * (2 * n)
        dp = ctypes.c_int(0)
        buf = np.empty((H, W1, Dp), np.uint16)
        self._check(self.lib.b200sgm_debug_read(self.h, lane, what.encode(), buf.ctypes.data_as(ctypes.c_void_p),
                                                ctypes.c_size_t(buf.nbytes), ctypes.byref(dp)))
        return buf[:, :, :p.numDisparities]

    def debug_image(self, what, W, H, lane=0) -> np.ndarray:
        buf = np.empty((H, W), np.int16)
        self._check(self.lib.b200sgm_debug_read(self.h, lane, what.encode(), buf.ctypes.data_as(ctypes.c_void_p),
                                                ctypes.c_size_t(buf.nbytes), None))
        return buf


def reproject_from_camera(Kl, Pl, Pr, depth_min, depth_max) -> CReproject:
    """b200sgm_reproject_from_camera: q and the disparity window formed exactly like the reference forms them."""
    lib = load_library()
    rp = CReproject()
    k, pl, pr = (np.ascontiguousarray(a, np.float64) for a in (Kl, Pl, Pr))
    lib.b200sgm_reproject_from_camera.restype = None
    lib.b200sgm_reproject_from_camera(ctypes.byref(rp), k.ctypes.data_as(ctypes.c_void_p), pl.ctypes.data_as(ctypes.c_void_p),
                                      pr.ctypes.data_as(ctypes.c_void_p), ctypes.c_double(depth_min), ctypes.c_double(depth_max))
    return rp


def alu_peak(device=0):
    """Measured packed-int16 issue peak in 1e12 elementary ops/s: (min3, min2, aggregation mix)."""
    lib = load_library()
    out = (ctypes.c_double * 3)()
    rc = lib.b200sgm_alu_peak(int(device), out)
    if rc != 0:
        raise B200SGMError(rc, "b200sgm_alu_peak failed")
    return tuple(out)
