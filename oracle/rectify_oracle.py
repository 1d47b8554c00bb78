"""CPU oracle (TEST INFRASTRUCTURE ONLY) for row N2 of SURVEY.md section 8(f): the rectification step in front of the
matcher, as the reference performs it in rectify() (/root/reference/src/generate_disparity.cpp:370-386 and
src/rectify.cpp:111-127):

    cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1, map1, map2);
    cv::remap(image, image_rect, map1, map2, cv::INTER_CUBIC, cv::BORDER_CONSTANT);

The arithmetic lives in OpenCV imgproc/calib3d (absent from /root/reference; oracle version = cv2 4.13.0).  This is a
numpy restatement of the published algorithm, pinned against cv2 itself by tests/test_rectify_oracle.py (live) and the
committed fixture tests/golden/rectify_small.npz (generator: tests/golden/make_rectify_golden.py).  Only tests/,
__graft_entry__.smoke() and bench.py's CPU leg may import this module; the product path never does.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS
COEF_BITS = 15


def _cubic_coeffs(x):
    """interpolateCubic (A = -0.75) in float32 arithmetic."""
    x = f32(x)
    A = f32(-0.75)
    c0 = ((A * (x + f32(1)) - f32(5) * A) * (x + f32(1)) + f32(8) * A) * (x + f32(1)) - f32(4) * A
    c1 = ((A + f32(2)) * x - (A + f32(3))) * x * x + f32(1)
    c2 = ((A + f32(2)) * (f32(1) - x) - (A + f32(3))) * (f32(1) - x) * (f32(1) - x) + f32(1)
    c3 = f32(1) - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], f32)


_TAB = None


def cubic_table() -> np.ndarray:
    """cv::remap's fixed-point bicubic table for 8-bit images: [1024][16] int16 weights, each set summing to 2**15.
    The sum is repaired on the largest / smallest weight found in rows and columns 2..3 of the 4x4 set (OpenCV's rule)."""
    global _TAB
    if _TAB is not None:
        return _TAB
    T = INTER_TAB_SIZE
    t1 = np.stack([_cubic_coeffs(f32(i) * f32(1.0 / T)) for i in range(T)])
    tab = np.zeros((T * T, 16), np.int64)
    for i in range(T):
        for j in range(T):
            v = (t1[i][:, None] * t1[j][None, :]).astype(f32)
            it = np.clip(np.rint((v * f32(1 << COEF_BITS)).astype(f32)).astype(np.int64), -32768, 32767)
            s = int(it.sum())
            if s != (1 << COEF_BITS):
                diff = s - (1 << COEF_BITS)
                Mk = mk = (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if it[k1, k2] < it[mk]:
                            mk = (k1, k2)
                        elif it[k1, k2] > it[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    it[Mk] -= diff
                else:
                    it[mk] -= diff
            tab[i * T + j] = it.reshape(16)
    _TAB = tab
    return tab


def remap_cubic(src: np.ndarray, map1: np.ndarray, map2: np.ndarray) -> np.ndarray:
    """cv::remap(src, map1, map2, INTER_CUBIC, BORDER_CONSTANT, 0) for CV_8UC1 and CV_32FC1 maps."""
    src = np.ascontiguousarray(src, np.uint8)
    H, W = src.shape
    sx = np.rint(map1.astype(f32) * f32(INTER_TAB_SIZE)).astype(np.int64)      # cvRound: half to even
    sy = np.rint(map2.astype(f32) * f32(INTER_TAB_SIZE)).astype(np.int64)
    wt = cubic_table()[(sy & 31) * 32 + (sx & 31)]
    ix = np.clip(sx >> INTER_BITS, -32768, 32767) - 1
    iy = np.clip(sy >> INTER_BITS, -32768, 32767) - 1
    S = src.astype(np.int64)
    acc = np.zeros(map1.shape, np.int64)
    for a in range(4):
        yy = iy + a
        oky = (yy >= 0) & (yy < H)
        for b in range(4):
            xx = ix + b
            ok = oky & (xx >= 0) & (xx < W)
            acc += np.where(ok, S[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)], 0) * wt[..., a * 4 + b]
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)


def inv3(m: np.ndarray) -> np.ndarray:
    """cv::invert on a 3x3 double matrix: cofactors over the determinant."""
    m = np.asarray(m, np.float64)
    det = m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0]) + \
        m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0])
    d = 1.0 / det
    t = np.empty((3, 3))
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def init_undistort_rectify_map(K, D, R, P, width: int, height: int):
    """cv::initUndistortRectifyMap(K, D, R, P, (width, height), CV_32FC1): double arithmetic per pixel, float32 maps.
    (OpenCV forms the homogeneous coordinates by repeated addition along the row; forming them directly changes a
    handful of map values per 5 Mpixel by one float32 ulp and none of the fixed-point coordinates remap derives.)"""
    K = np.asarray(K, np.float64).reshape(3, 3)
    R = np.eye(3) if R is None else np.asarray(R, np.float64).reshape(3, 3)
    P = np.asarray(P, np.float64).reshape(3, -1)[:, :3]
    k = np.zeros(14)
    if D is not None:
        d = np.asarray(D, np.float64).ravel()
        k[:d.size] = d
    k1, k2, p1, p2, k3, k4, k5, k6, s1, s2, s3, s4, tx, ty = k
    if tx != 0 or ty != 0:
        raise ValueError("tilted sensor model not restated")
    m = np.empty((3, 3))
    for r in range(3):
        for q in range(3):
            m[r, q] = P[r, 0] * R[0, q] + P[r, 1] * R[1, q] + P[r, 2] * R[2, q]
    ir = inv3(m)
    u0, v0, fx, fy = K[0, 2], K[1, 2], K[0, 0], K[1, 1]
    j = np.arange(width, dtype=np.float64)[None, :]
    i = np.arange(height, dtype=np.float64)[:, None]
    _x = j * ir[0, 0] + (i * ir[0, 1] + ir[0, 2])
    _y = j * ir[1, 0] + (i * ir[1, 1] + ir[1, 2])
    _w = j * ir[2, 0] + (i * ir[2, 1] + ir[2, 2])
    w = 1.0 / _w
    x = _x * w
    y = _y * w
    x2 = x * x
    y2 = y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2)
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2) + s1 * r2 + s2 * r2 * r2
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy + s3 * r2 + s4 * r2 * r2
    return (fx * xd + u0).astype(np.float32), (fy * yd + v0).astype(np.float32)


def rectify(image, K, D, R, P) -> np.ndarray:
    """rectify() of generate_disparity.cpp:370-386."""
    h, w = image.shape
    m1, m2 = init_undistort_rectify_map(K, D, R, P, w, h)
    return remap_cubic(image, m1, m2)


def sample_camera(width: int, height: int, seed: int = 0, strength: float = 1.0):
    """Synthetic camera model used by the tests and the bench (defined with the other input generators)."""
    import importlib
    return importlib.import_module("i3dr_stereo_camera-ros_b200.synth").sample_camera(width, height, seed, strength)
