"""Runs a few StereoBM frames at 2448x2048x256 (for ncu launch lists / captures of the k_bm_* kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import synth, Engine  # noqa: E402

W, H, D = 2448, 2048, 256
L, R = synth.make_pair(W, H, D, 0, 1000)
eng = Engine(0, W, H, D, 1)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    d = eng.bm_compute(L, R, D, 9, speckleWindowSize=100, speckleRange=2)
print("bm crc", synth.crc32(d), "valid %.1f%%" % (100 * (d != -16).mean()), "launches", eng.launch_count())
eng.close()
