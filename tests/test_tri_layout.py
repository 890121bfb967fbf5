"""Host logic of the one-CTA ILU(0) triangular solves (tuning key 4 = 2): the level-order layout built by the product header
navier-stokes-dealii_b200/csrc/nsg_tri_layout.h is walked on the CPU the way k_ilu_solve_cta walks it (tests/helpers/
tri_layout_check.cpp) - window, far reads, write-outs, ring capacity, the 8-lane summation order against the 32-lane order of the
level-scheduled kernels (bitwise) - and the result is compared with scipy's triangular solves."""
import ctypes as C
import importlib
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from conftest import mesh_path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ERRORS = {10: "the window no longer held a position when it was written out", 20: "a staged level does not fit the rings",
          21: "slot count of a level", 30: "position map", 31: "a row longer than its slabs", 32: "an entry twice or in the wrong slot",
          33: "a padding slot with a source", 34: "wrong column position", 35: "a column of the same or a later level",
          40: "a column flagged in-window lies outside the window", 41: "the window slot of a column was overwritten",
          42: "a far column had not landed in global memory", 43: "wrong unknown read", 50: "8-lane sum differs from the 32-lane order",
          60: "an entry without a slot"}
PRODUCTION = dict(window=8192, ring_slots=8192, ring_rows=1024, depth=2, flush=512)   # TRI_* of nsg_precond.cuh
SMALL = dict(window=256, ring_slots=512, ring_rows=64, depth=2, flush=16)              # far reads and write-outs on the small meshes
MEDIUM = dict(window=4096, ring_slots=2048, ring_rows=256, depth=2, flush=128)         # the same + levels read in place on cmy (levels of 838 rows)


@pytest.fixture(scope="module")
def helper():
    path = os.path.join(ROOT, "tests", "helpers", "libtri_layout_check.so")
    if not os.path.exists(path):
        pytest.skip("tests/helpers/libtri_layout_check.so not built (make)")
    L = C.CDLL(path)
    i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
    i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    L.tri_layout_check.restype = C.c_int
    L.tri_layout_check.argtypes = [C.c_int64, i64p, i32p, f64p, f64p, f64p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i64p]
    return L


def velocity_block(name, levels=0):
    """Pattern of the velocity block of the Jacobian on a repo mesh with diagonally dominant random values."""
    pkg = importlib.import_module("navier-stokes-dealii_b200")
    m = pkg.Mesh.read_msh(mesh_path(name))
    if levels:
        m = m.refine(levels)
    d = pkg.Dofs(m)
    part = pkg.Part(d, 0)
    n = d.n_u
    rp, cl = np.asarray(part.jac_rowptr), np.asarray(part.jac_col)
    A = sp.csr_matrix((np.ones(rp[n]), cl[: rp[n]], rp[: n + 1]), shape=(n, d.n))[:, :n].tocsr()
    A.sort_indices()
    rng = np.random.default_rng(5)
    A.data = rng.uniform(-1.0, 1.0, A.nnz)
    A.setdiag(8.0 + rng.random(n) + np.asarray(abs(A).sum(axis=1)).ravel())
    A.sort_indices()
    return A.tocsr()


def run(helper, A, limits):
    n = A.shape[0]
    x = np.random.default_rng(9).standard_normal(n)
    y, stats = np.zeros(n), np.zeros(8, np.int64)
    rc = helper.tri_layout_check(n, A.indptr.astype(np.int64), A.indices.astype(np.int32), np.ascontiguousarray(A.data, np.float64), x, y,
                                 limits["window"], limits["ring_slots"], limits["ring_rows"], limits["depth"], limits["flush"], stats)
    assert rc == 0, f"invariant {rc}: {'backward solve: ' if rc > 100 else ''}{ERRORS.get(rc % 100, '?')}"
    # Ifpack's apply: L unit lower, D the inverse pivots, U the strictly upper part scaled beforehand
    Lm = sp.tril(A, -1).tocsr() + sp.identity(n, format="csr")
    Um = sp.triu(A, 1).tocsr() + sp.identity(n, format="csr")
    yl = spla.spsolve_triangular(Lm, x, lower=True)
    want = spla.spsolve_triangular(Um, yl / A.diagonal(), lower=False)
    assert np.abs(y - want).max() <= 1e-12 * np.abs(want).max()
    return dict(zip(("levels_l", "levels_u", "slots_l", "slots_u", "far_reads", "in_place", "padding", "flushes"), (int(v) for v in stats)))


@pytest.mark.parametrize("name,levels", [("square_h0.1.msh", 0), ("square_h0.05.msh", 0), ("cylinder_cmy.msh", 0)])
def test_layout_walk_small_limits(helper, name, levels):
    """Small window and rings: columns outside the window, write-outs and levels read in place all occur on the small meshes."""
    st = run(helper, velocity_block(name, levels), MEDIUM if "cmy" in name else SMALL)
    assert st["far_reads"] > 0 and st["flushes"] > 2
    if "cmy" in name:
        assert st["in_place"] > 0


def test_layout_walk_refuses_a_window_too_small_for_the_levels(helper):
    """The kernel's precondition (widest level <= window / 4, TRI_MAX_LEVEL_ROWS): build_block does not offer the one-CTA solve then."""
    A = velocity_block("cylinder_cmy.msh")
    y, stats = np.zeros(A.shape[0]), np.zeros(8, np.int64)
    rc = helper.tri_layout_check(A.shape[0], A.indptr.astype(np.int64), A.indices.astype(np.int32), np.ascontiguousarray(A.data), y.copy(), y,
                                 256, 512, 64, 2, 16, stats)
    assert rc == -3


@pytest.mark.parametrize("name", ["square_h0.0125.msh", "cylinder_cmy.msh"])
def test_layout_walk_production_limits(helper, name):
    """The limits the kernel is built with, on the block of BASELINE.json configs[1] (51 842 rows, 1 120 + 1 120 levels) and on
    the reference's cylinder mesh (wide levels: some are read in place)."""
    st = run(helper, velocity_block(name), PRODUCTION)
    if "square" in name:
        assert st["levels_l"] == 1120 and st["levels_u"] == 1120 and st["in_place"] == 0 and st["far_reads"] > 0
    else:
        assert st["in_place"] > 0
