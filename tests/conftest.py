import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# tolerances of BASELINE.json north_star
E_RTOL = 1e-6      # energies: relative
F_RTOL = 1e-5      # forces: relative RMS (mixed precision)
E_RTOL_DISCARDED = 5e-5   # the partial energy returned when includeEnergy is false (FP32 pair terms)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def rel_rms(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = float((b ** 2).sum())
    return float(np.sqrt(((a - b) ** 2).sum() / den)) if den > 0 else float(np.abs(a - b).max())


@pytest.fixture(scope="session")
def build_native():
    """Make sure the CPU checkers are built (the CUDA library is built by __graft_entry__.build())."""
    import __graft_entry__ as g
    g.build_oracle()
    return True


def golden_case(name):
    import make_golden
    data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    pos, box, force = make_golden.build(name)
    return data, pos, box, force


GOLDEN_NAMES = ["c1_water64_nopbc", "water216_pbc", "fluxwater216_pbc", "water400_rect", "methanol_water_small", "rock_salt"]
