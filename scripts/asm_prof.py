"""Fixed assembly workload for ncu: level-L mesh, a few assemblies with the chosen variant / staging.
usage: python scripts/asm_prof.py L mesh variant stage"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
MESH = sys.argv[2] if len(sys.argv) > 2 else "cmy"
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 5
stage = int(sys.argv[4]) if len(sys.argv) > 4 else 0
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
dev = pkg.DeviceProblem(part, 0)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
dev.set_tuning(1, variant)
print("cells", m.n_cells, "N", d.n, "nnz", part.nnz_jac)
print("assembly ms", dev.time_kernel(0, 4))
dev.close()
