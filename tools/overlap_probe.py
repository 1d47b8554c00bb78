"""How long do stages of two different frames take when they share the GPU? (development aid, b200sgm_debug_overlap)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, Engine  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
eng = Engine(0, cfg.width, cfg.height, p.numDisparities, 2, p)
import numpy as np
outs = [np.empty(L.shape, np.int16) for _ in range(2)]
for lane in (0, 1):
    eng.enqueue(lane, L, R, outs[lane])
    eng.wait(lane)
lib = eng.lib
lib.b200sgm_debug_overlap.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.POINTER(ctypes.c_float)]
names = {1: "cost", 2: "horiz", 4: "vert", 3: "cost+horiz", 6: "horiz+vert", 7: "cost+horiz+vert", 0: "-"}
cases = [(1, 0), (2, 0), (4, 0), (4, 2), (2, 4), (4, 1), (1, 4), (2, 1), (1, 2), (4, 3), (6, 1), (7, 0), (7, 7)]
for a, b in cases:
    ms = ctypes.c_float(0)
    rc = lib.b200sgm_debug_overlap(eng.h, cfg.width, cfg.height, a, b, 5, ctypes.byref(ms))
    print("%-4s lane0: %-16s lane1: %-16s rc %d  %.3f ms" % (cfg.name, names[a], names[b], rc, ms.value), flush=True)
eng.close()
