"""Generates the golden vectors for the SGBM path from OpenCV itself (cv2 == the library the
reference's MatcherOpenCVSGBM calls), using the reference's call sequence (oracle/cv2_reference.py).

Run in the build container (needs cv2):   python tests/golden/make_golden.py
Writes tests/golden/sgbm_small.npz (small input/output pairs) and tests/golden/golden_crc.json
(CRC32 of inputs and cv2 outputs for the BASELINE configs c1, c2, c3, seed 1000).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import SGBMParams, CONFIGS, synth  # noqa: E402
from oracle import cv2_reference as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
          "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")

# (W, H, seed, extra_noise, params) -- covers minD in {0, 9, negative}, both modes, D not a multiple of 16,
# uniqueness 0, big penalties, speckle on/off, parameter-defaulting rules (<=0 values), even block size.
SMALL = [
    (96, 64, 1, 0, SGBMParams(numDisparities=32)),
    (96, 64, 2, 0, SGBMParams(numDisparities=32, mode=1, uniquenessRatio=10, disp12MaxDiff=1)),
    (130, 50, 3, 0, SGBMParams(numDisparities=48, minDisparity=9, blockSize=5)),
    (130, 50, 4, 12, SGBMParams(numDisparities=48, minDisparity=-8, blockSize=5, mode=1)),
    (110, 40, 5, 0, SGBMParams(numDisparities=16, minDisparity=-20, blockSize=3, speckleWindowSize=0)),
    (100, 48, 6, 20, SGBMParams(numDisparities=24, blockSize=15, P1=1800, P2=7200, preFilterCap=63)),
    (100, 48, 7, 0, SGBMParams(numDisparities=40, blockSize=8, P1=0, P2=0, preFilterCap=1, uniquenessRatio=-1, disp12MaxDiff=-1)),
    (120, 60, 8, 5, SGBMParams(numDisparities=64, blockSize=9, uniquenessRatio=0, disp12MaxDiff=5, speckleWindowSize=20, speckleRange=1)),
    (120, 60, 9, 5, SGBMParams(numDisparities=64, blockSize=9, uniquenessRatio=0, mode=1, speckleWindowSize=400, speckleRange=2)),
    (90, 30, 10, 0, SGBMParams(numDisparities=8, minDisparity=2, blockSize=7, P1=8, P2=32)),
    (72, 24, 11, 40, SGBMParams(numDisparities=16, minDisparity=1, blockSize=21, preFilterCap=7)),
    (160, 80, 12, 0, SGBMParams(numDisparities=128, blockSize=9, mode=1)),
]


def main():
    out = {}
    for i, (W, H, seed, noise, p) in enumerate(SMALL):
        L, R = synth.make_pair(W, H, p.numDisparities, p.minDisparity, seed)
        if noise:
            rng = np.random.default_rng(seed + 77)
            R = np.clip(R.astype(np.int32) + rng.integers(-noise, noise + 1, R.shape), 0, 255).astype(np.uint8)
        out["L%d" % i] = L
        out["R%d" % i] = R
        out["P%d" % i] = np.array([getattr(p, f) for f in FIELDS], np.int32)
        out["D%d" % i] = ref.compute(L, R, p)
    out["n"] = np.array(len(SMALL))
    np.savez_compressed(os.path.join(HERE, "sgbm_small.npz"), **out)

    crc = {}
    import cv2
    crc["opencv_version"] = cv2.__version__
    crc["numpy_version"] = np.__version__
    for name in ("c1", "c2", "c3", "cL"):
        c = CONFIGS[name]
        p = c.params
        L, R = synth.make_pair(c.width, c.height, p.numDisparities, p.minDisparity, 1000)
        d = ref.compute(L, R, p)
        crc[name] = dict(seed=1000, left=synth.crc32(L), right=synth.crc32(R), disp=synth.crc32(d),
                         valid_frac=float((d != p.invalid()).mean()), disp_sum=int(d.astype(np.int64).sum()))
        print(name, crc[name])
    with open(os.path.join(HERE, "golden_crc.json"), "w") as f:
        json.dump(crc, f, indent=1)


if __name__ == "__main__":
    main()
