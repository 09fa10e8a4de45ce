import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

LAYOUTS = {
    "418": ((418, 418), [(53, 53), (27, 27), (14, 14), (7, 7), (4, 4), (2, 2)]),
    "512": ((512, 512), [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]),
    "tiny": ((96, 96), [(6, 5), (3, 3), (2, 2), (1, 2), (1, 1), (1, 1)]),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_anchors(layout):
    """Reference anchors_all_layer output (list of [y,x,h,w]) from the fixture."""
    z = golden("anchors.npz")
    out = []
    for l in range(6):
        out.append([z["%s_l%d_%s" % (layout, l, k)] for k in "yxhw"])
    return out


def golden_files(prefix):
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
