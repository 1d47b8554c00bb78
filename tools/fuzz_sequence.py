"""One-off stress of ONE engine handle across changing image sizes, parameters, lanes and entry points (stale-state bugs),
plus random reprojection settings through b200sgm_compute_xyz.  python tools/fuzz_sequence.py [calls] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import SGBMParams, synth, Engine  # noqa: E402
from oracle import oracle  # noqa: E402

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 80
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
MAXW, MAXH, MAXD, LANES = 700, 420, 160, 3
eng = Engine(0, MAXW, MAXH, MAXD, LANES)
bad = 0
pending = {}
for it in range(n_calls):
    W = int(rng.integers(30, MAXW + 1)); H = int(rng.integers(4, MAXH + 1))
    D = int(rng.choice([16, 32, 48, 64, 80, 128, 144, 160])); minD = int(rng.choice([-16, 0, 0, 3, 9]))
    if W - (D + abs(minD)) < 4:
        continue
    p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=int(rng.choice([3, 5, 9, 15])), P1=int(rng.choice([200, 8, 600])),
                   P2=int(rng.choice([400, 32, 2400])), disp12MaxDiff=int(rng.choice([0, 1, 5])), preFilterCap=int(rng.choice([15, 31, 63])),
                   uniquenessRatio=int(rng.choice([0, 10, 15, 50])), speckleWindowSize=int(rng.choice([0, 100, 30])),
                   speckleRange=int(rng.choice([4, 2, 1])), mode=int(rng.integers(0, 2)))
    L, R = synth.make_pair(W, H, D, minD, seed=int(rng.integers(1 << 30)))
    want = oracle.compute(L, R, p)
    kind = int(rng.integers(0, 4))
    eng.set_params(p)
    if kind != 2 and 0 in pending:          # the synchronous entry points run on lane 0: it must be idle (else ESTATE "busy")
        eng.wait(0)
        o, w, tag = pending.pop(0)
        if not np.array_equal(o, w):
            bad += 1; print("MISMATCH (lane 0, deferred)", tag, int((o != w).sum()), flush=True)
    if kind == 0:
        got = eng.compute(L, R)
    elif kind == 1:
        got = np.round(eng.compute_f32(L, R)).astype(np.int16)
    elif kind == 2:
        ln = int(rng.integers(0, LANES))
        if ln in pending:
            eng.wait(ln)
            o, w, tag = pending.pop(ln)
            if not np.array_equal(o, w):
                bad += 1; print("MISMATCH (lane, deferred)", tag, int((o != w).sum()), flush=True)
        out = np.empty((H, W), np.int16)
        eng.enqueue(ln, L, R, out)
        pending[ln] = (out, want, (it, W, H, str(p)))
        continue
    else:
        fx = float(rng.uniform(300, 3000)); cx = W / 2 + float(rng.uniform(-5, 5)); cxr = cx + float(rng.choice([0.0, 0.0, 3.5, -2.25]))
        cy = H / 2 + float(rng.uniform(-5, 5)); base = float(rng.uniform(0.05, 0.5))
        q = oracle.calc_q(fx, cx, cxr, cy, -fx * base)
        depth_min = float(rng.choice([0.0, 0.3, 1.0])); depth_max = float(rng.choice([3.0, 10.0, 60.0]))
        fT = np.float32(base * fx)
        min_disp = float(fT / np.float32(depth_max))
        max_disp = float("inf") if depth_min == 0 else float(fT / np.float32(depth_min))
        got, dmat, depth, pts, n = eng.compute_xyz(L, R, q, depth_min, depth_max, min_disp, max_disp)
        wdm = oracle.process_disparity(want, min_disp, max_disp)
        wdepth, wpts = oracle.reproject(wdm, L, q, depth_min, depth_max)
        ok = np.array_equal(dmat, wdm) and np.array_equal(depth.view(np.uint32), wdepth.view(np.uint32)) and n == wpts.shape[0] and \
            np.array_equal(pts.view(np.uint32), wpts.view(np.uint32))
        if not ok:
            bad += 1; print("MISMATCH xyz", it, W, H, n, wpts.shape[0], depth_min, depth_max, flush=True)
    if not np.array_equal(got, want):
        bad += 1; print("MISMATCH kind %d it %d %dx%d %s: %d px" % (kind, it, W, H, p, int((got != want).sum())), flush=True)
for ln, (o, w, tag) in pending.items():
    eng.wait(ln)
    if not np.array_equal(o, w):
        bad += 1; print("MISMATCH (lane, final)", tag, int((o != w).sum()), flush=True)
eng.close()
print("sequence fuzz: %d calls, %d bad" % (n_calls, bad))
