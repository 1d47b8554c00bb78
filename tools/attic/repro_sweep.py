"""Development aid: replays tests/test_gpu_parity.py::test_random_parameter_sweep_vs_oracle, and for every mismatching case
prints where the pixels differ, whether the mismatch is deterministic, and what the hybrid path (set_path(2)) gives."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from b200sgm import SGBMParams, synth, Engine
from oracle import oracle

def run(L, R, p, path=0):
    H, W = L.shape
    eng = Engine(0, W, H, p.numDisparities, 1, p)
    eng.set_path(path)
    try:
        return eng.compute(L, R)
    finally:
        eng.close()

rng = np.random.default_rng(99)
nbad = 0
for it in range(60):
    W = int(rng.integers(70, 300)); H = int(rng.integers(24, 100))
    D = int(rng.choice([16, 32, 48, 8, 24, 40, 64, 80, 128, 144, 256])); minD = int(rng.choice([-8, 0, 1, 2, 9, -20, 30]))
    if W - (D + abs(minD)) < 8:
        continue
    p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=int(rng.choice([3, 5, 9, 15, 8, 21])),
                   P1=int(rng.choice([200, 8, 1800, 0])), P2=int(rng.choice([400, 32, 7200, 0])),
                   disp12MaxDiff=int(rng.choice([0, 1, 2, 5, -1])), preFilterCap=int(rng.choice([7, 31, 63, 1])),
                   uniquenessRatio=int(rng.choice([0, 2, 10, 15, -1, 50])), speckleWindowSize=int(rng.choice([0, 100, 20, 400])),
                   speckleRange=int(rng.choice([4, 2, 1, 0])), mode=int(rng.integers(0, 2)))
    L, R = synth.make_pair(W, H, D, minD, seed=int(rng.integers(1 << 30)))
    if it % 3 == 0:
        R = np.clip(R.astype(int) + rng.integers(-25, 26, R.shape), 0, 255).astype(np.uint8)
    want = oracle.compute(L, R, p)
    got = run(L, R, p)
    if not np.array_equal(got, want):
        nbad += 1
        ys, xs = np.nonzero(got != want)
        print("case", it, W, H, p, "W1", p.w1(W))
        print("  bad px", len(ys), list(zip(ys.tolist(), xs.tolist()))[:20])
        print("  got", got[ys, xs][:10], "want", want[ys, xs][:10])
        for k in range(3):
            g2 = run(L, R, p)
            print("  rerun", k, "same as first:", np.array_equal(g2, got), "bad:", int((g2 != want).sum()))
        gh = run(L, R, p, 2)
        print("  hybrid path bad:", int((gh != want).sum()))
        # the pre-filter stages: WTA output before median / speckle
        eng = Engine(0, W, H, D, 1, p); eng.compute(L, R)
        wta = eng.debug_image("wta", W, H); eng.close()
        _, st = oracle.compute(L, R, p, dumps=True)
        ys, xs = np.nonzero(wta != st["disp_wta"])
        print("  wta-stage bad px", len(ys), list(zip(ys.tolist(), xs.tolist()))[:20])
print("mismatching cases:", nbad)
