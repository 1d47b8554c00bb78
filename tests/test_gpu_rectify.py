"""GPU parity for row N2 (rectification in front of the matcher): b200sgm_set_camera / b200sgm_rectify through the C ABI
against the cv2 4.13 golden fixture and the numpy oracle.  The rectified image is integer work: bit-exact.  The CV_32FC1 maps
come from double arithmetic: within 1 float32 ulp of the oracle in at most a few pixels (stated tolerance)."""
import os

import numpy as np
import pytest

import b200sgm
from b200sgm import SGBMParams, synth, Engine
from oracle import oracle, rectify_oracle as ro

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ulp_diff(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


def test_rectify_matches_cv2_golden():
    z = np.load(os.path.join(ROOT, "tests", "golden", "rectify_small.npz"))
    eng = Engine(0, 320, 240, 16, 1)
    for n in range(int(z["n"])):
        g = {k: z["c%d_%s" % (n, k)] for k in ("img", "K", "D", "R", "P", "map1", "map2", "rect")}
        h, w = g["img"].shape
        eng.set_camera(n & 1, g["K"], g["D"], g["R"], g["P"])
        got = eng.rectify(n & 1, g["img"])
        assert np.array_equal(got, g["rect"]), "case %d: %d px differ from cv2" % (n, (got != g["rect"]).sum())
        m1, m2 = eng.rectify_maps(n & 1, w, h)
        for m, want in ((m1, g["map1"]), (m2, g["map2"])):
            d = _ulp_diff(m, want)
            assert d.max() <= 1 and (d != 0).sum() <= 4
    eng.close()


def test_rectify_random_cameras_vs_oracle():
    rng = np.random.default_rng(17)
    eng = Engine(0, 700, 500, 16, 1)
    for it in range(10):
        w, h = int(rng.integers(8, 700)), int(rng.integers(5, 500))
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        K, D, R, P = ro.sample_camera(w, h, 200 + it, float(rng.choice([0.0, 1.0, 5.0, 15.0])))
        if it % 3 == 1:
            D = np.concatenate([D, rng.uniform(-0.02, 0.02, 3), rng.uniform(-1e-3, 1e-3, 4)])
        if it == 4:
            R = None
        if it == 7:                      # a map that leaves the source image: constant border, partial footprints
            P = P.copy(); P[0, 2] += 0.6 * w; P[1, 2] -= 0.4 * h
        eng.set_camera(0, K, D, R, P)
        want = ro.rectify(img, K, D, R, P)
        got = eng.rectify(0, img)
        assert np.array_equal(got, want), "it %d (%dx%d): %d px differ" % (it, w, h, (got != want).sum())
        # strided input
        pad = np.zeros((h, w + 13), np.uint8); pad[:, :w] = img
        assert np.array_equal(eng.rectify(0, pad[:, :w]), want)
    eng.close()


def test_rectify_full_size_and_feeds_the_matcher():
    """2448x2048 (stereo_capture.launch:14-15): rectified pair == oracle, then straight into the matcher on the device."""
    import torch
    W, H, D = 2448, 2048, 64
    p = SGBMParams(numDisparities=D)
    L, R = synth.make_pair(W, H, D, 0, 31)
    cams = [ro.sample_camera(W, H, 7, 1.0), ro.sample_camera(W, H, 9, 1.0)]
    eng = Engine(0, W, H, D, 1, p)
    want = []
    for cam, (img, c) in enumerate(zip((L, R), cams)):
        eng.set_camera(cam, *c)
        want.append(ro.rectify(img, *c))
        assert np.array_equal(eng.rectify(cam, img), want[cam])
    dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
    rL, rR = torch.empty_like(dL), torch.empty_like(dR)
    dD = torch.empty((H, W), dtype=torch.int16, device="cuda")
    s = torch.cuda.Stream()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for rep in range(3):
        e0.record(s)
        eng.rectify_device(0, 0, dL.data_ptr(), W, W, H, rL.data_ptr(), W, stream=s.cuda_stream)
        eng.rectify_device(0, 1, dR.data_ptr(), W, W, H, rR.data_ptr(), W, stream=s.cuda_stream)
        e1.record(s)
        eng.compute_device(0, rL.data_ptr(), W, rR.data_ptr(), W, W, H, dD.data_ptr(), W * 2, stream=s.cuda_stream)
        e2.record(s)
        s.synchronize()
    assert np.array_equal(rL.cpu().numpy(), want[0]) and np.array_equal(rR.cpu().numpy(), want[1])
    got = dD.cpu().numpy()
    assert np.array_equal(got, eng.compute(want[0], want[1]))
    print("rectify both images %dx%d: %.3f ms; match: %.3f ms" % (W, H, e0.elapsed_time(e1), e1.elapsed_time(e2)))
    eng.close()


def test_rectify_error_behaviour():
    eng = Engine(0, 64, 64, 16, 1)
    img = np.zeros((32, 48), np.uint8)
    with pytest.raises(b200sgm.B200SGMError):       # no camera installed
        eng.rectify(0, img)
    K, D, R, P = ro.sample_camera(48, 32, 1)
    with pytest.raises(b200sgm.B200SGMError):       # camera index
        eng.set_camera(2, K, D, R, P)
    with pytest.raises(b200sgm.B200SGMError):       # tilted sensor model is not supported
        eng.set_camera(0, K, np.concatenate([D, np.zeros(7), [0.01, 0.0]]), R, P)
    with pytest.raises(b200sgm.B200SGMError):       # singular projection
        eng.set_camera(0, K, D, R, np.zeros((3, 4)))
    eng.set_camera(0, K, D, R, P)
    assert np.array_equal(eng.rectify(0, img), np.zeros_like(img))
    with pytest.raises(b200sgm.B200SGMError):       # exceeds the engine's max size
        eng.rectify(0, np.zeros((80, 48), np.uint8))
    eng.close()
