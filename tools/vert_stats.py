"""Counts how often an edge warp of k_vert found its neighbour's record late (B200SGM_DEBUG_VERT=4)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, Engine  # noqa: E402
cfg = CONFIGS["c3"]; p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
eng = Engine(0, cfg.width, cfg.height, p.numDisparities, 1, p)
for i in range(3):
    eng.compute(L, R)
buf = np.zeros(4, np.int32)
eng._check(eng.lib.b200sgm_debug_read(eng.h, 0, b"stats", buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(16), None))
print("late records over 3 frames:", buf[1], "of", 3 * 2 * 147 * cfg.height)
