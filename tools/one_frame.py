"""Runs a few frames of one config on one lane (for ncu launch lists / single-kernel captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, Engine  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
eng = Engine(0, cfg.width, cfg.height, p.numDisparities, 1, p)
eng.set_path(path)
for i in range(nframes):
    d = eng.compute(L, R)
print(cfg.name, "crc", synth.crc32(d), "launches", eng.launch_count())
eng.close()
