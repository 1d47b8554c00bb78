"""The C++ MatcherB200SGM adapter (the reference's AbstractStereoMatcher contract on top of the C ABI), driven by the
ROS-free harness that mirrors init_matcher / updateMatcher / stereo_match of generate_disparity.cpp."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from b200sgm import SGBMParams, synth
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    b = importlib.import_module("i3dr_stereo_camera-ros_b200.build")
    b.build()
    return b.build_host()


def test_harness_builds_and_prints_usage(harness):
    r = subprocess.run([harness], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


def run_harness(harness, tmp_path, L, R, p, full_dp=0, extra=()):
    H, W = L.shape
    lp, rp, op = tmp_path / "l.raw", tmp_path / "r.raw", tmp_path / "d.f32"
    L.tofile(lp); R.tofile(rp)
    if op.exists():
        op.unlink()
    args = [harness, str(lp), str(rp), str(W), str(H), str(op), p.minDisparity, p.numDisparities, p.blockSize, p.uniquenessRatio,
            p.speckleRange, p.speckleWindowSize, p.preFilterCap, p.P1, p.P2, full_dp] + list(extra)
    r = subprocess.run([str(a) for a in args], capture_output=True, text=True, timeout=300)
    return r, (np.fromfile(op, np.float32).reshape(H, W) if op.exists() else None)


@pytest.mark.gpu
def test_adapter_matches_oracle_through_node_call_sequence(harness, tmp_path):
    # the node's compiled-in defaults (generate_disparity.cpp:100-110): min_disparity 9, window 15, range 64
    for p, full_dp in ((SGBMParams(minDisparity=9, numDisparities=64, blockSize=15), 0),
                       (SGBMParams(minDisparity=0, numDisparities=128, blockSize=9, uniquenessRatio=10), 1)):
        L, R = synth.make_pair(400, 240, p.numDisparities, p.minDisparity, 21)
        r, got = run_harness(harness, tmp_path, L, R, p, full_dp)
        assert r.returncode == 0, r.stderr
        assert "Images MUST be the same resolution" in r.stderr      # the deliberate mismatched-size call at the end
        want = oracle.to_float(oracle.compute(L, R, p.replace(mode=full_dp)))
        assert got.dtype == np.float32 and np.array_equal(got, want)   # CV_32F holding disparity x16


@pytest.mark.gpu
def test_adapter_reports_bad_parameters_like_the_reference(harness, tmp_path):
    p = SGBMParams(numDisparities=4096)   # does not fit CV_16S x16 -> forwardMatch returns -1, message on stderr
    L, R = synth.make_pair(200, 60, 64, 0, 3)
    r, got = run_harness(harness, tmp_path, L, R, p)
    assert r.returncode == 1
    assert "Error in B200 SGM parameters" in r.stderr


@pytest.mark.gpu
def test_block_matcher_adapter_through_node_call_sequence(harness, tmp_path):
    """MatcherB200BM (row N4) behind the same init_matcher / updateMatcher / stereo_match sequence: CV_32F holding the x16
    value of cv::StereoBM with the node's parameters (window 15, range 64, texture 10, uniqueness 15, speckle 100/4)."""
    from oracle import bm_oracle as bo
    p = SGBMParams(minDisparity=0, numDisparities=64, blockSize=15, uniquenessRatio=15, speckleWindowSize=100, speckleRange=4,
                   preFilterCap=31)
    L, R = synth.make_pair(400, 240, 64, 0, 22)
    r, got = run_harness(harness, tmp_path, L, R, p, 0, extra=(1, 10))
    assert r.returncode == 0, r.stderr
    want = bo.compute(L, R, 64, 15, 0, 31, 10, 15, 100, 4).astype(np.float32)
    assert got.dtype == np.float32 and np.array_equal(got, want)
    # a window that does not fit the image: OpenCV throws, the reference returns -1 and prints its message
    r, got = run_harness(harness, tmp_path, L[:12], R[:12], p, 0, extra=(1, 10))
    assert r.returncode == 1 and "Error in OpenCV StereoBM parameters" in r.stderr


@pytest.mark.gpu
def test_adapter_right_view_disparity_and_fused_cloud(harness, tmp_path):
    """backwardMatch(): the right-view disparity of cv::ximgproc::createRightMatcher (matcherOpenCVSGBM.cpp:46-51) = the same
    matcher with minDisparity -(minD + D) + 1, uniqueness 0, disp12MaxDiff 1000000 and no speckle filter on (right, left) --
    pinned through cv2 when importable.  matchToCloud(): rows a11 + R fused behind the adapter, BGR8 colour."""
    from oracle import cv2_reference as ref
    p = SGBMParams(minDisparity=3, numDisparities=64, blockSize=9)
    W, H = 400, 240
    L, R = synth.make_pair(W, H, 64, 3, 33)
    rng = np.random.default_rng(9)
    bgr = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
    cp = tmp_path / "c.raw"; bgr.tofile(cp)
    fx, cx, cxr, cy, p14, dmin, dmax = 600.0, 200.0, 203.5, 120.0, -600.0 * 0.12, 0.4, 37.3
    extra = ["--back", tmp_path / "b.f32", "--cloud", tmp_path / "cl", fx, cx, cxr, cy, p14, dmin, dmax, cp, 3]
    r, got = run_harness(harness, tmp_path, L, R, p, 0, [0, 10] + extra)
    assert r.returncode == 0, r.stderr
    want = oracle.compute(L, R, p)
    assert np.array_equal(got, oracle.to_float(want))
    # right view
    pr = p.replace(minDisparity=-(3 + 64) + 1, uniquenessRatio=0, disp12MaxDiff=1000000, speckleWindowSize=0, speckleRange=0)
    back = np.fromfile(tmp_path / "b.f32", np.float32).reshape(H, W)
    assert np.array_equal(back, oracle.to_float(oracle.compute(R, L, pr)))
    if ref.have_cv2():
        import cv2
        m = cv2.StereoSGBM_create(-(3 + 64) + 1, 64, 9)
        m.setUniquenessRatio(0); m.setP1(p.P1); m.setP2(p.P2); m.setPreFilterCap(p.preFilterCap)
        m.setDisp12MaxDiff(1000000); m.setSpeckleWindowSize(0)
        assert np.array_equal(back, m.compute(R, L).astype(np.float32))
    # fused cloud
    Kl = np.array([[fx, 0, cx], [0, fx, cy], [0, 0, 1.0]])
    q = oracle.calc_q(fx, cx, cxr, cy, p14)
    lo, hi = oracle.disparity_window(fx, -p14 / fx, dmin, dmax)
    wdm = oracle.process_disparity(want, lo, hi)
    wdepth, wpts = oracle.reproject(wdm, bgr, q, dmin, dmax)
    dmat = np.fromfile(str(tmp_path / "cl") + ".dmat.f32", np.float32).reshape(H, W)
    depth = np.fromfile(str(tmp_path / "cl") + ".depth.f32", np.float32).reshape(H, W)
    raw = np.fromfile(str(tmp_path / "cl") + ".cloud.bin", np.uint8)
    n = int(raw[:4].view(np.uint32)[0])
    pts = raw[4:].view(np.float32).reshape(-1, 4)
    assert np.array_equal(dmat, wdm) and np.array_equal(depth, wdepth)
    assert n == wpts.shape[0] and n > 1000 and np.array_equal(pts.view(np.uint32), wpts.view(np.uint32))


@pytest.mark.gpu
def test_block_matcher_adapter_rejects_bad_prefilter_size(harness, tmp_path):
    """cv::StereoBM::compute throws for an even or out-of-range preFilterSize although PREFILTER_XSOBEL never reads it, and the
    node pushes that value (generate_disparity.cpp:255): MatcherB200BM must fail the same way (forwardMatch == -1)."""
    p = SGBMParams(numDisparities=64, blockSize=9)
    L, R = synth.make_pair(300, 100, 64, 0, 4)
    r, got = run_harness(harness, tmp_path, L, R, p, 0, [1, 10])
    assert r.returncode == 0, r.stderr
    for bad in (8, 3, 257):
        r, _ = run_harness(harness, tmp_path, L, R, p, 0, [1, 10, "--bad-prefilter-size", bad])
        assert r.returncode == 1 and "preFilterSize" in r.stderr


@pytest.mark.gpu
def test_cpp_frame_stream_driver(tmp_path):
    """host/stream_driver.cpp (frame i -> GPU i mod N, one engine + worker thread per GPU, `lanes` frames in flight, no
    torch, no NCCL) through host/stream_bench on the GPUs of this box: every distinct frame's disparity CRC equals the oracle's,
    whatever GPU and lane produced it."""
    import json
    import torch
    b = importlib.import_module("i3dr_stereo_camera-ros_b200.build")
    b.build()
    exe = b.build_stream()
    p = SGBMParams(numDisparities=64, minDisparity=2)
    W, H, distinct = 320, 200, 5
    pairs = [synth.make_pair(W, H, 64, 2, 70 + i) for i in range(distinct)]
    raw = tmp_path / "pairs.raw"
    with open(raw, "wb") as f:
        for L, R in pairs:
            f.write(L.tobytes()); f.write(R.tobytes())
    want = [synth.crc32(oracle.compute(L, R, p)) for L, R in pairs]
    for gpus in sorted({1, torch.cuda.device_count()}):
        for lanes in (1, 3):
            args = [exe, raw, distinct, W, H, 23, gpus, lanes, p.minDisparity, p.numDisparities, p.blockSize, p.uniquenessRatio, p.speckleRange,
                    p.speckleWindowSize, p.preFilterCap, p.P1, p.P2, p.mode]
            r = subprocess.run([str(a) for a in args], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr
            out = json.loads(r.stdout)
            assert out["crc"] == want and out["status"] == 0 and out["frames"] == 23, out
