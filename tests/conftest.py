import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """build (or reuse) the CUDA library and the oracle once per session"""
    from vren_b200 import build

    build.build_cuda()
    build.build_oracle()
    build.build_reference_extract()
    return True


@pytest.fixture(scope="session")
def vren(built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vren_b200 import lib

    lib.load()
    return lib


def splitmix64(seed: int, n: int):
    """counter-based PRNG: same stream on every platform"""
    import numpy as np

    with np.errstate(over="ignore"):
        z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) + np.uint64(seed) * np.uint64(0xD1B54A32D192ED03)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z
