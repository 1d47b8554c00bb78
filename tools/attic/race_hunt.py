"""Development aid: repeats small cases around the non-deterministic mismatch and counts bad runs per variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from b200sgm import SGBMParams, synth, Engine
from oracle import oracle

def trial(W, H, p, reps=6, seed=5):
    L, R = synth.make_pair(W, H, p.numDisparities, p.minDisparity, seed)
    want = oracle.compute(L, R, p)
    eng = Engine(0, W, H, p.numDisparities, 1, p)
    bad = []
    for k in range(reps):
        got = eng.compute(L, R)
        bad.append(int((got != want).sum()))
    eng.close()
    return bad

base = dict(minDisparity=0, numDisparities=144, blockSize=15, P1=8, P2=400, preFilterCap=7, uniquenessRatio=2, speckleWindowSize=0)
for W, H, kw in [(175, 56, dict(mode=1)), (175, 56, dict(mode=0)), (175, 200, dict(mode=1)), (175, 200, dict(mode=0)),
                 (206, 56, dict(mode=1)), (159, 56, dict(mode=1, numDisparities=128)), (159, 56, dict(mode=0, numDisparities=128)),
                 (150, 56, dict(mode=1)), (150, 56, dict(mode=0)), (160, 56, dict(mode=1, numDisparities=64)), (300, 120, dict(mode=1)),
                 (95, 56, dict(mode=1, numDisparities=64)), (95, 56, dict(mode=0, numDisparities=64)), (79, 90, dict(mode=1, numDisparities=48)),
                 (79, 90, dict(mode=0, numDisparities=48))]:
    p = SGBMParams(**{**base, **kw})
    print(W, H, "W1", p.w1(W), "D", p.numDisparities, "mode", p.mode, "bad px per run:", trial(W, H, p), flush=True)
