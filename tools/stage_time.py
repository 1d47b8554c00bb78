"""Per-stage CUDA-event timings of one config on one lane, frames back to back (development aid)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json
import b200sgm  # noqa: E402
from b200sgm import CONFIGS, synth, Engine  # noqa: E402

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
nframes = int(sys.argv[2]) if len(sys.argv) > 2 else 6
path = int(sys.argv[3]) if len(sys.argv) > 3 else 0
p = cfg.params
L, R = synth.make_pair(cfg.width, cfg.height, p.numDisparities, p.minDisparity, 1000)
eng = Engine(0, cfg.width, cfg.height, p.numDisparities, 1, p)
eng.set_path(path)
try:
    d = eng.compute(L, R)
except Exception as e:
    print('first frame error:', e)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_crc.json")))
key = "c3" if cfg.name in ("c3", "c4", "c5") else cfg.name
eng.profile(True)
eng.stage_times(0)
for i in range(nframes):
    try:
        d = eng.compute(L, R)
    except Exception as e:
        pass
ms, n = eng.stage_times(0)
tot = sum(ms.values()) / n
print(cfg.name, "path", path, "crc_ok", synth.crc32(d) == gold[key]["disp"], "frame_ms %.3f" % tot,
      " ".join("%s=%.3f" % (k, v / n) for k, v in ms.items()), flush=True)
eng.close()
