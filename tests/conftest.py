import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def gold():
    import numpy as np
    return {name: np.load(os.path.join(GOLD, name + ".npz")) for name in ("unet_simple", "unet_openai", "operators", "sampler_tiny", "simplified")}
