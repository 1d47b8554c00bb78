"""TEST INFRASTRUCTURE: the reference's SGBM call sequence on OpenCV's own binding (cv2).

cv::StereoSGBM is the third-party code the reference's MatcherOpenCVSGBM delegates to
(/root/reference/src/stereoMatcher/matcherOpenCVSGBM.cpp:14,21).  This module drives the
same library through `cv2` with exactly the reference's sequence:
  create(64, 9, 5)                                  matcherOpenCVSGBM.cpp:14
  setNumDisparities, setBlockSize, setMinDisparity, setUniquenessRatio, setSpeckleRange,
  setSpeckleWindowSize, setPreFilterCap, setP1, setP2      generate_disparity.cpp:245-256
  compute(left, right)                              matcherOpenCVSGBM.cpp:21
disp12MaxDiff and mode are NOT forwarded by the reference; they are set here only when a
config asks for them (BASELINE config 2).  Used by tests/, golden generation and bench.py's
cpu_baseline / --impl reference legs only.
"""
from __future__ import annotations

import numpy as np


def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def make_matcher(p):
    import cv2
    m = cv2.StereoSGBM_create(64, 9, 5)
    m.setNumDisparities(int(p.numDisparities))
    m.setBlockSize(int(p.blockSize))
    m.setMinDisparity(int(p.minDisparity))
    m.setUniquenessRatio(int(p.uniquenessRatio))
    m.setSpeckleRange(int(p.speckleRange))
    m.setSpeckleWindowSize(int(p.speckleWindowSize))
    m.setPreFilterCap(int(p.preFilterCap))
    m.setP1(int(p.P1))
    m.setP2(int(p.P2))
    if p.disp12MaxDiff != 0:
        m.setDisp12MaxDiff(int(p.disp12MaxDiff))
    if p.mode == 1:
        m.setMode(cv2.STEREO_SGBM_MODE_HH)
    return m


def compute(left: np.ndarray, right: np.ndarray, p) -> np.ndarray:
    return make_matcher(p).compute(left, right)
