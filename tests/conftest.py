import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    from b200sgm import SGBMParams
    z = np.load(os.path.join(ROOT, "tests", "golden", "sgbm_small.npz"))
    fields = ("minDisparity", "numDisparities", "blockSize", "P1", "P2", "disp12MaxDiff", "preFilterCap",
              "uniquenessRatio", "speckleWindowSize", "speckleRange", "mode")
    cases = []
    for i in range(int(z["n"])):
        p = SGBMParams(**{f: int(v) for f, v in zip(fields, z["P%d" % i])})
        cases.append((z["L%d" % i], z["R%d" % i], p, z["D%d" % i]))
    return cases


@pytest.fixture(scope="session")
def golden_crc():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_crc.json")) as f:
        return json.load(f)
