import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name))


def golden_sd(g):
    import torch
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd/")}


def rel_l2(a, b):
    import torch
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    return float((a - b).abs().pow(2).sum().sqrt() / b.abs().pow(2).sum().sqrt().clamp_min(1e-30))


@pytest.fixture(scope="session")
def ref_ns():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box): verbatim-reference checks run in the build container")
    return ref_loader.load()
