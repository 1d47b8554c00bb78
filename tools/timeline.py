"""Multi-lane stage timeline (development aid): which stages of which frames overlap on the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, Engine, STAGES  # noqa: E402

cfg = CONFIGS["c3"]
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
p = cfg.params
W, H = cfg.width, cfg.height
L, R = synth.make_pair(W, H, p.numDisparities, 0, 1000)
eng = Engine(0, W, H, p.numDisparities, lanes, p)
dL, dR = torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()
dD = torch.empty((lanes, H, W), dtype=torch.int16, device="cuda")
for i in range(lanes * 2):
    eng.compute_device(i % lanes, dL.data_ptr(), W, dR.data_ptr(), W, W, H, dD[i % lanes].data_ptr(), W * 2)
torch.cuda.synchronize()
eng.profile(True)
for i in range(nfr):
    eng.compute_device(i % lanes, dL.data_ptr(), W, dR.data_ptr(), W, W, H, dD[i % lanes].data_ptr(), W * 2)
torch.cuda.synchronize()
ev = []
for ln in range(lanes):
    eng.stage_times(ln)
    tl = eng.stage_timeline(ln)
    for f, row in enumerate(tl):
        for s, name in enumerate(STAGES):
            ev.append((row[s], row[s + 1], ln, f, name))
ev.sort()
t0 = ev[0][0]
for a, b_, ln, f, name in ev:
    if name in ("cost", "horizontal", "vertical_wta", "speckle"):
        print("%8.3f -> %8.3f  (%6.3f)  lane %d frame %d %s" % (a - t0, b_ - t0, b_ - a, ln, f, name))
print("total %.3f ms for %d frames -> %.1f fps" % (max(e[1] for e in ev) - t0, nfr, nfr / (max(e[1] for e in ev) - t0) * 1e3))
