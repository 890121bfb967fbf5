import importlib
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


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("navier-stokes-dealii_b200")


@pytest.fixture(scope="session")
def golden():
    return GOLDEN


def mesh_path(name):
    return os.path.join(GOLDEN, name)


# (file, surface entity) -> survey §8d facts: V, T, E, n_u, n_p, nnz(J), nnz(Stokes), nnz(Mp)
MESH_FACTS = {
    ("cylinder_cmy.msh", -1): dict(V=3350, T=6448, E=9798, n_u=26296, n_p=3350, nnz=(867738, 844792, 22946)),
    ("square_h0.0125.msh", -1): dict(V=6561, T=12800, E=19360, n_u=51842, n_p=6561, nnz=(1717609, 1672328, 45281)),
    ("cylinder_mesh2d.msh", 5): dict(V=175, T=288, E=463, n_u=1276, n_p=175, nnz=None),
}


def analytic_state(dofs, scale=1.0):
    """SURVEY §8d throughput state: u = (sin(pi x)cos(pi y), -cos(pi x)sin(pi y)), p = x y."""
    xy = dofs.support_points()
    n_u = dofs.n_u
    sol = np.zeros(dofs.n)
    sol[0:n_u:2] = np.sin(np.pi * xy[0:n_u:2, 0]) * np.cos(np.pi * xy[0:n_u:2, 1])
    sol[1:n_u:2] = -np.cos(np.pi * xy[1:n_u:2, 0]) * np.sin(np.pi * xy[1:n_u:2, 1])
    sol[n_u:] = xy[n_u:, 0] * xy[n_u:, 1]
    return scale * sol


def row_scaled_err(a, b, rowptr):
    """max |a-b| relative to the largest entry of the same row (entries are sums of cancelling
    cell contributions, so an entry-wise relative error is not meaningful)."""
    a, b = np.asarray(a), np.asarray(b)
    n = len(rowptr) - 1
    lens = np.diff(rowptr)
    rows = np.repeat(np.arange(n), lens)
    scale = np.zeros(n)
    np.maximum.at(scale, rows, np.abs(b))
    scale[scale == 0] = 1.0
    return float((np.abs(a - b) / scale[rows]).max()) if len(a) else 0.0
