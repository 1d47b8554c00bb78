"""Rectifies a few 2448x2048 frames on the device (for ncu captures of k_rectify_maps / k_remap_cubic)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import synth, Engine  # noqa: E402

W, H = 2448, 2048
L, R = synth.make_pair(W, H, 64, 0, 1000)
eng = Engine(0, W, H, 64, 1, b200sgm.SGBMParams(numDisparities=64))
for cam in (0, 1):
    eng.set_camera(cam, *synth.sample_camera(W, H, 7 + 2 * cam, 1.0))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    a, b = eng.rectify(0, L), eng.rectify(1, R)
print("rectified crc", synth.crc32(a), synth.crc32(b), "launches", eng.launch_count())
eng.close()
