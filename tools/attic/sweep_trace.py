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
buf = np.zeros(32 * 32 * 4 + 2 * 320, np.int64)
eng._check(eng.lib.b200sgm_debug_read(eng.h, 0, b"trace", buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes), None))
W1 = p.w1(cfg.width)
tw = -(-W1 // min(148, W1 // 2))
t = buf[:4096].reshape(32, 32, 4)
be = buf[4096:].reshape(-1, 2)
dur = (be[:, 1] - be[:, 0])[be[:, 0] > 0]
print("per-strip sweep cycles: min %d median %d max %d (strip %d)" % (dur.min(), np.median(dur), dur.max(), int(dur.argmax())))
print("slowest 8 strips:", np.argsort(dur)[-8:], dur[np.argsort(dur)[-8:]])
if int(row0) == -2:
    seg = buf[:32 * 4].reshape(32, 4)
    print("cycles per segment over the whole sweep (path: 0 barrier+loop, 1 prefetch+wait+loads, 2 V/B steps+publish, 3 A step+stage hand-over;")
    print(" agent: 0 store+arrive+loop, 1 poll, 2 slot wait)")
    for w in range(2 * tw + 1):
        if seg[w].sum() > 0:
            print("w%02d" % w, " ".join("%10d" % x for x in seg[w]), " total %d" % seg[w].sum())
    eng.close(); sys.exit(0)
if int(row0) < 0:
    rows = buf[:2048]; ag = buf[2048:4096]
    n = int((rows > 0).sum())
    d = np.diff(rows[:n])
    print("path warp 0 row period: median %d mean %d p90 %d p99 %d max %d cycles over %d rows" % (np.median(d), d.mean(), np.percentile(d, 90), np.percentile(d, 99), d.max(), n))
    big = np.nonzero(d > 4 * np.median(d))[0]
    print("rows with period > 4x median:", len(big), "share of time %.2f" % (d[big].sum() / d.sum()), "first:", big[:40])
    for lo in range(0, n - 1, 128):
        print("rows %4d-%4d mean period %6d" % (lo, min(lo + 127, n - 1), d[lo:lo + 128].mean()))
    eng.close(); sys.exit(0)
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
