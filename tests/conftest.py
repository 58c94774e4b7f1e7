import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def model():
    from trex_gym_b200.model_compiler import load_builtin

    return load_builtin()


@pytest.fixture(scope="session")
def action_limits(model):
    lo = model["mb_lower"][1:][model["obs_dof"]]
    hi = model["mb_upper"][1:][model["obs_dof"]]
    return lo, hi


def rel_err(ref, got):
    """max |ref - got| / max(|ref|, floor) per state block -- the per-step parity measure."""
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    return float(np.abs(ref - got).max() / max(np.abs(ref).max(), 1e-3))


STATE_BLOCKS = {"pos": slice(0, 3), "quat": slice(3, 7), "omega": slice(7, 10), "vel": slice(10, 13),
                "q": slice(13, 38), "qd": slice(38, 63), "tau": slice(63, 88)}
