"""Generates tests/golden/rectify_small.npz: inputs and cv2 4.13 outputs of the reference's rectification call sequence
(cv2.initUndistortRectifyMap(K, D, R, P, size, CV_32FC1) + cv2.remap(INTER_CUBIC, BORDER_CONSTANT)), the oracle pin
for row N2.  Run in the build container (cv2 importable): python tests/golden/make_rectify_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import rectify_oracle as ro  # noqa: E402

out = {}
cases = [(160, 120, 0, 1.0), (200, 96, 1, 3.0), (97, 131, 2, 8.0), (320, 240, 3, 1.0)]
for n, (w, h, seed, strength) in enumerate(cases):
    rng = np.random.default_rng(100 + seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (5, 5), 1.1)
    K, D, R, P = ro.sample_camera(w, h, seed, strength)
    m1, m2 = cv2.initUndistortRectifyMap(K, D, R, P, (w, h), cv2.CV_32FC1)
    rect = cv2.remap(img, m1, m2, cv2.INTER_CUBIC, borderMode=cv2.BORDER_CONSTANT)
    for k, v in dict(img=img, K=K, D=D, R=R, P=P, map1=m1, map2=m2, rect=rect).items():
        out["c%d_%s" % (n, k)] = v
out["n"] = np.int64(len(cases))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rectify_small.npz"), **out)
print("wrote", len(cases), "cases, cv2", cv2.__version__)
