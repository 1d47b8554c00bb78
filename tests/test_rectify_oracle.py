"""Oracle pin for row N2 (rectification): the numpy restatement in oracle/rectify_oracle.py against the committed cv2 4.13
fixture and, when cv2 is importable, against cv2 live on further random cameras.  CPU only."""
import os
import zlib

import numpy as np
import pytest

from oracle import rectify_oracle as ro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "rectify_small.npz"))
    return [{k: z["c%d_%s" % (n, k)] for k in ("img", "K", "D", "R", "P", "map1", "map2", "rect")} for n in range(int(z["n"]))]


def test_cubic_table_invariants():
    t = ro.cubic_table()
    assert t.shape == (1024, 16) and (t.sum(axis=1) == 32768).all()
    # integer position: 1.0 saturates to 32767 as int16 and the missing unit lands on tap (2, 2) -- OpenCV's quirk, kept
    assert list(t[0]) == [0] * 5 + [32767] + [0] * 4 + [1] + [0] * 5
    assert zlib.crc32(t.astype(np.int16).tobytes()) == 0x690308d7


def test_remap_matches_cv2_golden(golden):
    for n, g in enumerate(golden):
        got = ro.remap_cubic(g["img"], g["map1"], g["map2"])
        assert np.array_equal(got, g["rect"]), "case %d: %d px differ" % (n, (got != g["rect"]).sum())


def test_maps_and_rectify_match_cv2_golden(golden):
    for n, g in enumerate(golden):
        h, w = g["img"].shape
        m1, m2 = ro.init_undistort_rectify_map(g["K"], g["D"], g["R"], g["P"], w, h)
        for m, want in ((m1, g["map1"]), (m2, g["map2"])):
            d = _ulp_diff(m, want)
            assert d.max() <= 1 and (d != 0).sum() <= 4, "case %d: %d map values off, max %d ulp" % (n, (d != 0).sum(), d.max())
        assert np.array_equal(ro.rectify(g["img"], g["K"], g["D"], g["R"], g["P"]), g["rect"])


def test_live_cv2_random_cameras():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for it in range(8):
        w, h = int(rng.integers(40, 400)), int(rng.integers(30, 300))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        K, D, R, P = ro.sample_camera(w, h, 50 + it, float(rng.choice([0.0, 1.0, 4.0, 12.0])))
        if it == 3:
            D = np.concatenate([D, rng.uniform(-0.02, 0.02, 3)])                      # rational model k4..k6
        if it == 5:
            D = np.concatenate([D, rng.uniform(-0.02, 0.02, 3), rng.uniform(-1e-3, 1e-3, 4)])   # + thin prism s1..s4
        if it == 6:
            R = None
        m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
        want = cv2.remap(img, m1, m2, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
        a1, a2 = ro.init_undistort_rectify_map(K, D, R, P, w, h)
        assert _ulp_diff(a1, m1).max() <= 1 and _ulp_diff(a2, m2).max() <= 1
        assert np.array_equal(ro.remap_cubic(img, m1, m2), want)
        assert np.array_equal(ro.rectify(img, K, D, R, P), want)
