"""Pins the CPU restatement (oracle/sgbm_oracle.c) against OpenCV: committed golden vectors (generated
from cv2 by tests/golden/make_golden.py) and, when cv2 is importable, live randomized comparisons."""
import numpy as np
import pytest

from b200sgm import SGBMParams, CONFIGS, synth
from oracle import oracle, cv2_reference as ref


def test_oracle_matches_golden_small(golden_small):
    for i, (L, R, p, want) in enumerate(golden_small):
        got = oracle.compute(L, R, p)
        assert np.array_equal(got, want), "golden case %d: %d px differ (%s)" % (i, (got != want).sum(), p)


def test_synth_and_oracle_match_golden_crc_c1(golden_crc):
    c = CONFIGS["c1"]
    L, R = synth.make_pair(c.width, c.height, c.params.numDisparities, c.params.minDisparity, 1000)
    g = golden_crc["c1"]
    assert synth.crc32(L) == g["left"] and synth.crc32(R) == g["right"]
    d = oracle.compute(L, R, c.params)
    assert synth.crc32(d) == g["disp"]
    assert int(d.astype(np.int64).sum()) == g["disp_sum"]


def test_synth_crc_c2_c3(golden_crc):
    for name in ("c2", "c3"):
        c = CONFIGS[name]
        L, R = synth.make_pair(c.width, c.height, c.params.numDisparities, c.params.minDisparity, 1000)
        assert synth.crc32(L) == golden_crc[name]["left"] and synth.crc32(R) == golden_crc[name]["right"]


@pytest.mark.skipif(not ref.have_cv2(), reason="cv2 not importable")
def test_oracle_matches_cv2_random():
    rng = np.random.default_rng(20261018)
    n = 0
    for it in range(40):
        W = int(rng.integers(70, 131)); H = int(rng.integers(24, 71))
        D = int(rng.choice([16, 32, 48, 8, 24, 40])); minD = int(rng.choice([-8, 0, 1, 2, 9, -20]))
        if W - (D + abs(minD)) < 8:
            continue
        p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=int(rng.choice([3, 5, 9, 15, 8])),
                       P1=int(rng.choice([200, 8, 1800, 0])), P2=int(rng.choice([400, 32, 7200, 0])),
                       disp12MaxDiff=int(rng.choice([0, 1, 2, 5, -1])), preFilterCap=int(rng.choice([7, 31, 63, 1])),
                       uniquenessRatio=int(rng.choice([0, 2, 10, 15, -1])), speckleWindowSize=int(rng.choice([0, 100, 20, 400])),
                       speckleRange=int(rng.choice([4, 2, 1, 0])), mode=int(rng.integers(0, 2)))
        L, R = synth.make_pair(W, H, D, minD, seed=int(rng.integers(1 << 30)))
        if it % 3 == 0:
            R = np.clip(R.astype(int) + rng.integers(-20, 21, R.shape), 0, 255).astype(np.uint8)
        want = ref.compute(L, R, p)
        got = oracle.compute(L, R, p)
        assert np.array_equal(got, want), "%d px differ for %s (%dx%d)" % ((got != want).sum(), p, W, H)
        n += 1
    assert n >= 25


@pytest.mark.skipif(not ref.have_cv2(), reason="cv2 not importable")
def test_postfilters_match_cv2():
    import cv2
    rng = np.random.default_rng(5)
    img = (rng.integers(-1, 40, (60, 90)) * 16).astype(np.int16)
    img[rng.random(img.shape) < 0.2] = -16
    assert np.array_equal(oracle.median3x3(img), cv2.medianBlur(img, 3))
    for size, rng_ in ((100, 4), (10, 1), (400, 2)):
        want = img.copy()
        cv2.filterSpeckles(want, -16, size, 16 * rng_)
        assert np.array_equal(oracle.filter_speckles(img, -16, size, 16 * rng_), want)


def test_degenerate_width_all_invalid():
    p = SGBMParams(numDisparities=64)
    L = np.zeros((8, 40), np.uint8)
    d = oracle.compute(L, L, p)
    assert (d == p.invalid()).all()


def test_reprojection_restatement_matches_numpy():
    """Row R (disparity_to_depth.cpp:136-205): the C restatement against an independent numpy float32 statement."""
    rng = np.random.default_rng(3)
    H, W = 40, 64
    disp16 = (rng.integers(-1, 60, (H, W)) * 16 + rng.integers(0, 16, (H, W))).astype(np.int16)
    q = oracle.calc_q(2400.0, 1224.0, 1224.0, 1024.0, -720.0)
    fT = np.float32(0.3 * 2400.0)
    dmat = oracle.process_disparity(disp16, float(fT / np.float32(10.0)), float("inf"))
    d = disp16.astype(np.float32) / np.float32(16)
    d[d < fT / np.float32(10.0)] = 10000
    assert np.array_equal(dmat, d)
    gray = rng.integers(0, 256, (H, W)).astype(np.uint8)
    depth, pts = oracle.reproject(dmat, gray, q, 0.0, 10.0)
    w = d * q[3] + q[4]
    with np.errstate(divide="ignore", invalid="ignore"):
        Z = q[2] / w
        X = (np.arange(W, dtype=np.float32)[None, :] + q[0]) / w
        Y = (np.arange(H, dtype=np.float32)[:, None] + q[1]) / w
    keep = (d != 0) & (d != 10000) & (w > 0) & (Z > 0) & (Z <= 10.0) & (Z >= 0.0)
    assert np.array_equal(depth, np.where(keep, Z, 0).astype(np.float32))
    assert pts.shape[0] == keep.sum()
    assert np.array_equal(pts[:, 0], X[keep]) and np.array_equal(pts[:, 1], Y[keep]) and np.array_equal(pts[:, 2], Z[keep])


def _camera(rng):
    fx = float(rng.uniform(600, 3000)); fy = fx * float(rng.uniform(0.99, 1.01))
    cx, cy = float(rng.uniform(200, 1300)), float(rng.uniform(150, 1100))
    cxr = cx + float(rng.choice([0.0, 0.0, 3.25, -7.5]))
    fp = fx * float(rng.uniform(0.95, 1.05))          # P's focal length differs from K's after rectification
    base = float(rng.uniform(0.05, 0.6))
    Kl = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    Pl = np.array([[fp, 0, cx, 0], [0, fp, cy, 0], [0, 0, 1.0, 0]])
    Pr = np.array([[fp, 0, cxr, -fp * base], [0, fp, cy, 0], [0, 0, 1.0, 0]])
    return Kl, Pl, Pr, fp, base


@pytest.mark.skipif(oracle.ref_lib() is None, reason="oracle/_ref not built (needs /root/reference or a prebuilt library)")
def test_reprojection_pinned_to_the_reference_code():
    """Rows a11 and R: the oracle's restatement against the REFERENCE'S OWN statements (oracle/_ref: calc_q and the per-pixel loop of
    disparity_to_depth.cpp, the float conversion and depth window of generate_disparity.cpp:436-452, extracted from /root/reference
    at build time and compiled against stand-in cv::Mat / pcl types).  Bit-for-bit, MONO8 and BGR8, several cameras and windows."""
    rng = np.random.default_rng(17)
    for case in range(12):
        H, W = int(rng.integers(20, 70)), int(rng.integers(30, 100))
        Kl, Pl, Pr, fp, base = _camera(rng)
        depth_min = float(rng.choice([0.0, 0.3, 1.0])); depth_max = float(rng.choice([10.0, 2.5, 100.0, 0.7]))
        disp16 = (rng.integers(-1, 200, (H, W)) * 16 + rng.integers(0, 16, (H, W))).astype(np.int16)
        disp16[rng.random((H, W)) < 0.15] = -16
        disp16[rng.random((H, W)) < 0.05] = 0
        # a10 + a11: what stereo_match() hands to processDisparity is CV_32F holding the x16 value
        want_dmat, lo, hi = oracle.ref_process_disparity(oracle.to_float(disp16), fp, base, depth_min, depth_max)
        assert (lo, hi) == oracle.disparity_window(fp, base, depth_min, depth_max)
        got_dmat = oracle.process_disparity(disp16, lo, hi)
        assert np.array_equal(got_dmat.view(np.uint32), want_dmat.view(np.uint32))
        # calc_q in double, cast to float as disparity_to_depth.cpp:136-140 does
        Q = oracle.ref_calc_q(Kl, Pr, Pl)
        q = oracle.calc_q(Kl[0, 0], Pl[0, 2], Pr[0, 2], Pl[1, 2], Pr[0, 3])
        assert np.array_equal(q, np.array([Q[0, 3], Q[1, 3], Q[2, 3], Q[3, 2], Q[3, 3]]).astype(np.float32))
        for color in (rng.integers(0, 256, (H, W)).astype(np.uint8), rng.integers(0, 256, (H, W, 3)).astype(np.uint8)):
            wd, wp = oracle.ref_reproject(want_dmat, color, Kl, Pr, Pl, depth_min, depth_max)
            gd, gp = oracle.reproject(got_dmat, color, q, depth_min, depth_max)
            assert np.array_equal(gd.view(np.uint32), wd.view(np.uint32))
            assert gp.shape == wp.shape and np.array_equal(gp.view(np.uint32), wp.view(np.uint32))
