"""One-off randomized parity sweep of the StereoBM path against the numpy oracle: python tools/fuzz_bm.py [cases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import synth, Engine  # noqa: E402
from oracle import bm_oracle as bo  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
eng = Engine(0, 900, 400, 256, 1)
bad = done = 0
while done < n_cases:
    W, H = int(rng.integers(24, 900)), int(rng.integers(8, 400))
    nd = int(rng.choice([16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256])); bs = int(rng.choice([5, 7, 9, 11, 15, 21, 31, 51]))
    mind = int(rng.choice([0, 0, 0, 4, 9, -3, -20, -64, 30, 147]))
    if bs >= min(W, H) or W * H * nd > 30e6:
        continue
    cap = int(rng.choice([1, 7, 15, 31, 63])); tex = int(rng.choice([0, 10, 100, 300, 2000, 20000])); uniq = int(rng.choice([0, 1, 5, 15, 40, 99, 150]))
    sw = int(rng.choice([0, 0, 10, 50, 200])); sr = int(rng.choice([0, 1, 2, 4, 32, 200]))
    L, R = synth.make_pair(W, H, nd, mind, int(rng.integers(1 << 30)))
    k = int(rng.integers(0, 4))
    if k == 1:
        R = np.clip(R.astype(int) + rng.integers(-20, 21, R.shape), 0, 255).astype(np.uint8)
    elif k == 2:
        L = rng.integers(0, 256, L.shape).astype(np.uint8); R = rng.integers(0, 256, R.shape).astype(np.uint8)
    elif k == 3:
        L = (L // 64 * 64).astype(np.uint8); R = (R // 64 * 64).astype(np.uint8)
    want = bo.compute(L, R, nd, bs, mind, cap, tex, uniq, sw, sr)
    try:
        got = eng.bm_compute(L, R, nd, bs, mind, cap, tex, uniq, sw, sr)
    except Exception as e:
        print("EXC", (W, H, nd, bs, mind, cap, tex, uniq, sw, sr), e, flush=True); bad += 1; done += 1; continue
    nb = int((got != want).sum())
    if nb:
        bad += 1; print("MISMATCH %d px" % nb, (W, H, nd, bs, mind, cap, tex, uniq, sw, sr), "kind", k, flush=True)
    done += 1
eng.close()
print("bm fuzz: %d cases, %d bad" % (done, bad))
