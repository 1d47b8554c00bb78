"""One-off randomized parity sweep of the CUDA path against the CPU oracle (larger and more varied than the test-suite's
60 cases; run on a GPU box: python tools/fuzz_parity.py [cases] [seed] [wide|tall|full])."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import SGBMParams, synth, Engine  # noqa: E402
from oracle import oracle  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2024)
bad = done = 0
t0 = time.time()
while done < n_cases:
    if len(sys.argv) > 3 and sys.argv[3] == "full":     # Phobos resolution
        W, H = 2448, 2048
    elif len(sys.argv) > 3 and sys.argv[3] == "tall":
        W = int(rng.integers(60, 400)); H = int(rng.integers(600, 2100))
    elif len(sys.argv) > 3 and sys.argv[3] == "wide":     # full-width strips (14-16 columns per vertical-sweep CTA), few rows
        W = int(rng.integers(1500, 2600)); H = int(rng.integers(3, 48))
    else:
        W = int(rng.integers(40, 900)); H = int(rng.integers(8, 320))
    D = int(rng.choice([16, 32, 48, 8, 24, 40, 64, 80, 96, 112, 128, 144, 160, 192, 256, 272])); minD = int(rng.choice([-64, -8, 0, 0, 0, 1, 2, 9, -20, 30, 147]))
    if W - (D + abs(minD)) < 4:
        continue
    p = SGBMParams(minDisparity=minD, numDisparities=D, blockSize=int(rng.choice([1, 3, 5, 7, 9, 9, 11, 15, 8, 21, 25])),
                   P1=int(rng.choice([200, 8, 1800, 0, 24, 600])), P2=int(rng.choice([400, 32, 7200, 0, 96, 2400])),
                   disp12MaxDiff=int(rng.choice([0, 1, 2, 5, -1, 100])), preFilterCap=int(rng.choice([7, 31, 63, 1, 15, 100])),
                   uniquenessRatio=int(rng.choice([0, 2, 10, 15, -1, 50, 99])), speckleWindowSize=int(rng.choice([0, 100, 20, 400, 5])),
                   speckleRange=int(rng.choice([4, 2, 1, 0, 8])), mode=int(rng.integers(0, 2)))
    L, R = synth.make_pair(W, H, D, minD, seed=int(rng.integers(1 << 30)))
    k = int(rng.integers(0, 4))
    if k == 1:
        R = np.clip(R.astype(int) + rng.integers(-25, 26, R.shape), 0, 255).astype(np.uint8)
    elif k == 2:
        L = rng.integers(0, 256, L.shape).astype(np.uint8); R = rng.integers(0, 256, R.shape).astype(np.uint8)
    elif k == 3:
        L = (L // 32 * 32).astype(np.uint8); R = (R // 32 * 32).astype(np.uint8)     # flat regions: many cost ties
    # stay inside the int16 cost contract of SURVEY 8c
    bs = (p.blockSize if p.blockSize > 0 else 5) | 1
    ft = max(p.preFilterCap, 15) | 1
    if bs * bs * (2 * ft + 63) + max(p.P2, 5) > 32767:
        continue
    try:
        want = oracle.compute(L, R, p)
        eng = Engine(0, W, H, D, 1, p)
        got = eng.compute(L, R)
        eng.close()
    except Exception as e:
        print("EXC", W, H, p, e, flush=True)
        bad += 1; done += 1
        continue
    nb = int((got != want).sum())
    if nb:
        bad += 1
        print("MISMATCH %d px: %dx%d kind %d uniq %d mode %d D %d bs %d %s" % (nb, W, H, k, p.uniquenessRatio, p.mode, p.numDisparities, p.blockSize, p), flush=True)
    done += 1
print("fuzz: %d cases, %d bad, %.0f s" % (done, bad, time.time() - t0))
