"""Oracle pin for row N4 (StereoBM): oracle/bm_oracle.py against the committed cv2 4.13 fixture and, when cv2 is importable,
against cv2 live on random parameters.  CPU only."""
import os

import numpy as np
import pytest

from oracle import bm_oracle as bo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))




def test_bm_matches_cv2_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "bm_small.npz"))
    for n in range(int(z["n"])):
        nd, bs, mind, cap, tex, uniq, sw, sr = (int(v) for v in z["c%d_p" % n])
        got = bo.compute(z["c%d_L" % n], z["c%d_R" % n], nd, bs, mind, cap, tex, uniq, sw, sr)
        want = z["c%d_disp" % n]
        assert np.array_equal(got, want), "case %d: %d px differ" % (n, (got != want).sum())


def test_bm_live_cv2_random():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    done = 0
    for trial in range(60):
        W, H = int(rng.integers(40, 200)), int(rng.integers(24, 90))
        nd = int(rng.choice([16, 32, 48, 64])); bs = int(rng.choice([5, 7, 9, 15, 21])); mind = int(rng.choice([0, 0, 4, -3, -20, 30, 60]))
        if bs >= min(W, H) or W - bs // 2 <= max(0, mind + nd - 1) + bs // 2:
            continue          # empty valid ROI: cv2 returns uninitialised memory there
        cap = int(rng.choice([1, 15, 31, 63])); tex = int(rng.choice([0, 10, 300, 2000])); uniq = int(rng.choice([0, 5, 15, 40]))
        sw = int(rng.choice([0, 0, 50, 200])); sr = int(rng.choice([0, 1, 4, 32]))
        T = cv2.GaussianBlur(rng.integers(0, 256, (H, W + 120)).astype(np.uint8), (5, 5), 1.0)
        s = int(rng.integers(0, 20))
        L = T[:, 60:60 + W].copy(); R = T[:, 60 + s:60 + s + W].copy()
        if trial % 3 == 0:
            R = np.clip(R.astype(int) + rng.integers(-6, 7, R.shape), 0, 255).astype(np.uint8)
        if trial % 7 == 3:
            L = (L // 64 * 64).astype(np.uint8); R = (R // 64 * 64).astype(np.uint8)
        m = cv2.StereoBM_create(64, 9)
        m.setMinDisparity(mind); m.setNumDisparities(nd); m.setBlockSize(bs); m.setPreFilterCap(cap); m.setTextureThreshold(tex)
        m.setUniquenessRatio(uniq); m.setSpeckleWindowSize(sw); m.setSpeckleRange(sr)
        want = m.compute(L, R)
        got = bo.compute(L, R, nd, bs, mind, cap, tex, uniq, sw, sr)
        assert np.array_equal(got, want), "trial %d: %d px differ" % (trial, (got != want).sum())
        done += 1
    assert done >= 20
