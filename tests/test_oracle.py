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
