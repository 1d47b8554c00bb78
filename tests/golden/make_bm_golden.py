"""Generates tests/golden/bm_small.npz: inputs and cv2 4.13 outputs of the reference's StereoBM call sequence
(cv2.StereoBM_create(64, 9) + the setters of matcherOpenCVBlock.cpp:52-110), the oracle pin for row N4.
Run in the build container (cv2 importable): python tests/golden/make_bm_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import synth  # noqa: E402

cases = [  # W, H, numDisparities, blockSize, minDisparity, preFilterCap, textureThreshold, uniquenessRatio, speckleWindow, speckleRange
    (160, 120, 64, 9, 0, 31, 10, 15, 0, 0), (200, 90, 32, 15, -8, 15, 100, 5, 60, 4), (131, 77, 16, 5, 9, 63, 0, 0, 0, 0),
    (240, 100, 48, 21, 0, 31, 10, 15, 100, 32), (96, 64, 16, 7, -20, 7, 500, 40, 20, 1),
]
out = {"n": np.int64(len(cases))}
for n, (W, H, nd, bs, mind, cap, tex, uniq, sw, sr) in enumerate(cases):
    L, R = synth.make_pair(W, H, nd, mind, 300 + n)
    m = cv2.StereoBM_create(64, 9)
    m.setMinDisparity(mind); m.setNumDisparities(nd); m.setBlockSize(bs); m.setPreFilterCap(cap); m.setTextureThreshold(tex)
    m.setUniquenessRatio(uniq); m.setSpeckleWindowSize(sw); m.setSpeckleRange(sr)
    out["c%d_L" % n] = L; out["c%d_R" % n] = R
    out["c%d_p" % n] = np.array([nd, bs, mind, cap, tex, uniq, sw, sr], np.int64)
    out["c%d_disp" % n] = m.compute(L, R)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bm_small.npz"), **out)
print("wrote", len(cases), "cases, cv2", cv2.__version__)
