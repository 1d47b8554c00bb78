"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; run on the GPU box)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import SGBMParams, CONFIGS, synth, Engine  # noqa: E402
from oracle import oracle  # noqa: E402


def diag(W, H, p, seed=1, path=0, verbose=True):
    L, R = synth.make_pair(W, H, p.numDisparities, p.minDisparity, seed)
    want, st = oracle.compute(L, R, p, dumps=True)
    eng = Engine(0, W, H, p.numDisparities, 1, p)
    eng.set_path(path)
    got = eng.compute(L, R)
    res = {}
    if p.w1(W) > 0:
        C = eng.debug_volume("C", W, H)
        res["C"] = int((C != st["C"].view(np.uint16)).sum())
        try:
            S = eng.debug_volume("S", W, H)
            res["S"] = int((S != st["S"].view(np.uint16)).sum())
        except Exception as e:  # fused paths may not materialise S
            res["S"] = str(e)
    res["wta"] = int((eng.debug_image("wta", W, H) != st["disp_wta"]).sum())
    res["median"] = int((eng.debug_image("median", W, H) != st["disp_med"]).sum())
    res["final"] = int((got != want).sum())
    if verbose:
        print("%4dx%-4d %s -> %s" % (W, H, p, res), flush=True)
        if res.get("C"):
            bad = np.argwhere(C != st["C"].view(np.uint16))
            print("   first C mismatches (y,x1,k):", bad[:5].tolist(), "got", [int(C[tuple(b)]) for b in bad[:5]],
                  "want", [int(st["C"].view(np.uint16)[tuple(b)]) for b in bad[:5]])
        elif res.get("S") and isinstance(res["S"], int):
            bad = np.argwhere(S != st["S"].view(np.uint16))
            print("   first S mismatches (y,x1,k):", bad[:5].tolist(), "got", [int(S[tuple(b)]) for b in bad[:5]],
                  "want", [int(st["S"].view(np.uint16)[tuple(b)]) for b in bad[:5]])
        elif res["wta"]:
            bad = np.argwhere(eng.debug_image("wta", W, H) != st["disp_wta"])
            print("   first wta mismatches (y,x):", bad[:8].tolist())
    eng.close()
    return res


if __name__ == "__main__":
    path = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    cases = [
        (96, 64, SGBMParams(numDisparities=32)),
        (96, 64, SGBMParams(numDisparities=32, mode=1)),
        (130, 50, SGBMParams(numDisparities=48, minDisparity=9, blockSize=5)),
        (130, 50, SGBMParams(numDisparities=48, minDisparity=-8, blockSize=5, mode=1)),
        (100, 48, SGBMParams(numDisparities=24, blockSize=15, P1=1800, P2=7200, preFilterCap=63)),
        (120, 60, SGBMParams(numDisparities=64, uniquenessRatio=0, disp12MaxDiff=5, speckleWindowSize=20, speckleRange=1)),
        (90, 30, SGBMParams(numDisparities=8, minDisparity=2, blockSize=7, P1=8, P2=32)),
        (300, 120, SGBMParams(numDisparities=128, mode=1)),
        (400, 100, SGBMParams(numDisparities=256)),
        (640, 480, CONFIGS["c1"].params),
    ]
    tot = 0
    for W, H, p in cases:
        t = time.time()
        r = diag(W, H, p, path=path)
        tot += r["final"]
    print("TOTAL final mismatches:", tot)
