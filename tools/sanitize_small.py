"""Small invocation of every kernel family (for compute-sanitizer --tool memcheck / racecheck): SGBM fast path in both modes,
StereoBM, rectification, reprojection."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import SGBMParams, synth, Engine  # noqa: E402

W, H, D = 333, 97, 64
L, R = synth.make_pair(W, H, D, 0, 5)
eng = Engine(0, W, H, 128, 2)
for p in (SGBMParams(numDisparities=D), SGBMParams(numDisparities=48, minDisparity=-8, blockSize=5, mode=1),
          SGBMParams(numDisparities=128, blockSize=15, preFilterCap=127)):
    eng.set_params(p)
    d = eng.compute(L, R)
    print("sgbm", p.numDisparities, p.mode, synth.crc32(d))
print("bm", synth.crc32(eng.bm_compute(L, R, 64, 9, speckleWindowSize=50, speckleRange=4)), synth.crc32(eng.bm_compute(L, R, 32, 21, minDisparity=9)))
for cam in (0, 1):
    eng.set_camera(cam, *synth.sample_camera(W, H, 3 + cam, 6.0))
print("rect", synth.crc32(eng.rectify(0, L)), synth.crc32(eng.rectify(1, R)))
q = np.array([-W / 2, -H / 2, 300.0, 1 / 0.1, 0.0], np.float32)
eng.set_params(SGBMParams(numDisparities=D))
out = eng.compute_xyz(L, R, q, 0.0, 50.0, 0.5, float("inf"))
print("xyz points", out[4])
eng.close()
