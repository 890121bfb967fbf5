"""ncu target: a few default-variant SpMVs on the bench workload (mesh, levels as in bench.py)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
MESH = sys.argv[1] if len(sys.argv) > 1 else "mesh2d"
L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
dev = pkg.DeviceProblem(part, 0)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
dev.assemble()
dev.set_delta(np.random.default_rng(0).standard_normal(part.n_own))
nb = 12 * part.nnz_jac + 8 * part.n_loc + 8 * part.n_own + 8 * (part.n_own + 1)
x = np.random.default_rng(1).standard_normal(part.n_own)
ref = None
for v in [int(a) for a in (sys.argv[3] if len(sys.argv) > 3 else "4").split(",")]:
    dev.set_tuning(0, v)
    y = dev.spmv(x) if part.n_own < 40_000_000 else None
    if ref is None:
        ref = y
    same = None if y is None else bool(np.array_equal(y, ref))
    dev.time_kernel(1, 2)
    ms = dev.time_kernel(1, 6)
    print("variant", v, "cells", m.n_cells, "N", d.n, "nnz", part.nnz_jac, "algorithmic bytes", nb, "ms", ms, "GB/s", nb / ms / 1e6,
          "bitwise equal to the first variant:", same, flush=True)
dev.close()
