"""Development aid: per-warp clock64 stamps of one strip of k_sweep over 32 rows (B200SGM_TRACE="strip,row0").
Path warps: 0 row start, 1 inputs ready, 2 states published, 3 row done.  WTA warps: 0 batch start, 1 rows parked, 2 batch done.
Agent: 0 slot free, 1 records arrived."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
strip, row0 = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("70", "1000")
os.environ["B200SGM_TRACE"] = "%s,%s" % (strip, row0)
import ctypes
import numpy as np
from b200sgm import CONFIGS, synth, Engine
cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
eng = Engine(0, cfg.width, cfg.height, p.numDisparities, 1, p)
eng.compute(L, R); eng.compute(L, R)
buf = np.zeros(32 * 32 * 4, np.int64)
eng._check(eng.lib.b200sgm_debug_read(eng.h, 0, b"trace", buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes), None))
t = buf.reshape(32, 32, 4)
t0 = t[t > 0].min()
W1 = p.w1(cfg.width)
tw = -(-W1 // min(148, W1 // 2))
np.set_printoptions(linewidth=250)
print("rows %s.., strip %s, tw %d; cycles since first stamp" % (row0, strip, tw))
for w in range(2 * tw + 1):
    role = "path" if w < tw else ("wta" if w < 2 * tw else "agent")
    for k in range(4):
        v = t[w, :12, k]
        if (v > 0).any():
            print("%5s w%02d p%d" % (role, w, k), " ".join("%7d" % (x - t0 if x > 0 else -1) for x in v))
eng.close()
