import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: statistical / long-running checks")


class Golden:
    """One of the reference's knitr caches (tests/golden/*.npz, see make_golden.py)."""

    def __init__(self, name):
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
        s = lambda k: self.z["setting." + k]
        self.times = s("times")
        self.x_r = s("x_r")
        self.x_i = s("x_i")
        self.t_correct = float(s("t_correct")[0])
        self.gridsize = int(s("gridsize")[0])
        self.init = {k: s("Init." + k) for k in ("C", "D", "y", "gridrep", "ng", "t", "l")}
        self.final = {k[len("final."):]: self.z[k] for k in self.z.files if k.startswith("final.")}
        self.iters = {k[len("iter."):]: self.z[k] for k in self.z.files if k.startswith("iter.")}
        # read eagerly: NpzFile decompresses on every access and is not thread-safe (the long-horizon tests replay the
        # oracle chains from a thread pool)
        self._settings = {k[len("setting."):]: self.z[k] for k in self.z.files if k.startswith("setting.")}

    def setting(self, key):
        return self._settings[key]


@pytest.fixture(scope="session")
def changepoint():
    return Golden("changepoint_sim")


@pytest.fixture(scope="session")
def constant():
    return Golden("constant_sim")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0
