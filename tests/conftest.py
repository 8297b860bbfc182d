import ctypes as C
import pathlib
import subprocess
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

P = 0xFFFFFFFF00000001


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def splitmix64_stream(seed: int, count: int) -> np.ndarray:
    """SplitMix64 values reduced mod p (BASELINE.md §3 synthetic input generator)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z % np.uint64(P)


def random_columns(ncols: int, n: int, seed: int = 0x5EED000000000000) -> np.ndarray:
    return np.stack([splitmix64_stream(seed | c, n) for c in range(ncols)]).astype(np.uint64)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/liborc.so), built on demand.  Test infrastructure only."""
    from oracle import binding
    return binding.load()


@pytest.fixture(scope="session")
def zkm():
    """The product library initialised on cuda:0 (GPU tests only)."""
    from zkm_b200 import lib as zl
    return zl.init(0)
