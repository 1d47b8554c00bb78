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
