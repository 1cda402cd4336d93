import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIXTURES = ["small_ra_slam_problem", "single_rpm", "single_range"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built here if missing; nvcc cross-compiles on CPU)."""
    from cora_b200 import build, capi
    build.build()
    return capi.load()


def load_fixture(name):
    from oracle import cora_oracle as co
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    p = co.parse_pyfg(str(g["pyfg"]), from_text=True)
    return g, p


def load_dataset(name, rank=None, preconditioner=1):
    """plaza2 / single_drone measurement arrays -> oracle Problem (Q assembled)."""
    from oracle import cora_oracle as co
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    arrays = {k: g[k] for k in g.files if k not in ("d", "n", "l")}
    p = co.Problem.from_arrays(int(g["d"]), int(g["n"]), int(g["l"]), arrays, rank=rank,
                               preconditioner=preconditioner)
    return p


def make_handle(p, preconditioner=1, **kw):
    from cora_b200 import capi
    return capi.Handle(p.d, p.n, p.m, p.n + p.l, p.Q, preconditioner=preconditioner, **kw)
