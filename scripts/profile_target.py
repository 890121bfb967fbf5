"""Small fixed workload for ncu: level-L CMY mesh, a few assemblies and SpMVs (both variants)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MESH = sys.argv[2] if len(sys.argv) > 2 else "cmy"
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
dev = pkg.DeviceProblem(part, 0)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
dev.set_delta(np.random.default_rng(0).standard_normal(part.n_own))
print("cells", m.n_cells, "N", d.n, "nnz", part.nnz_jac)
nb = 12 * part.nnz_jac + 8 * part.n_loc + 16 * part.n_own
for rep in range(2):
    for av in (0, 1):
        dev.set_tuning(1, av)
        print("assembly variant", av, "ms", dev.time_kernel(0, 3))
    for v in (1, 2, 4, 6):
        dev.set_tuning(0, v)
        ms = dev.time_kernel(1, 5)
        print("spmv variant", v, "ms", ms, "GB/s", nb / ms / 1e6)
    print("add_and_dot ms", dev.time_kernel(2, 5), "dot ms", dev.time_kernel(3, 5))
dev.apply_dirichlet(ld, lv)
print("dirichlet ms", dev.phase_ms())
dev.close()
