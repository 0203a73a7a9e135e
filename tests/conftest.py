import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def trpcage():
    from oracle import binding as ob
    xyz = ob.read_xtc(os.path.join(DATA, "trp-cage.xtc"))
    mass = ob.read_masses(os.path.join(DATA, "trp-cage.pdb"))
    return xyz, mass


def load_pts(name, dim):
    import numpy as np
    pts = np.fromfile(os.path.join(DATA, name), dtype=np.float64)
    return pts[: (pts.size // dim) * dim].reshape(-1, dim)
