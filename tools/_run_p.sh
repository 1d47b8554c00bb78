python - <<'P'
import sys, os, subprocess, numpy as np
sys.path.insert(0, os.getcwd())
import b200sgm
from b200sgm import CONFIGS, synth
cfg = CONFIGS["c3"]; p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
L.tofile("/tmp/l.raw"); R.tofile("/tmp/r.raw")
h = "i3dr_stereo_camera-ros_b200/host/harness"
args = [h, "/tmp/l.raw", "/tmp/r.raw", cfg.width, cfg.height, "/tmp/d.f32", p.minDisparity, p.numDisparities, p.blockSize, p.uniquenessRatio, p.speckleRange, p.speckleWindowSize, p.preFilterCap, p.P1, p.P2, p.mode, 0, 10, "--bench", 20]
r = subprocess.run([str(a) for a in args], capture_output=True, text=True)
print(r.stderr[-400:])
P
