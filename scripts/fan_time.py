"""Timing only: assembly variant 5 on a refined mesh. usage: fan_time.py L mesh [pf]"""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]); MESH = sys.argv[2]
pfs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, MESH, L, 1, 0)
dev = pkg.DeviceProblem(part, 0)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
ref = None
for pipe in (0,):
    for pf in pfs:
        dev.set_tuning(6, pf)
        dev.time_kernel(0, 3)
        ms = dev.time_kernel(0, 10)
        J, R = np.concatenate([dev.get_matrix_values(), dev.get_pm_values()]), dev.get_residual()
        if ref is None:
            ref = (J, R)
        same = np.array_equal(J, ref[0]) and np.array_equal(R, ref[1])
        print(f"{os.environ.get('NSG_LIB', 'default lib')}: {MESH} L{L} cells {m.n_cells} pipelined {pipe} pf {pf & 0xffff}/{pf >> 16}: {ms:8.3f} ms "
              f"{d.n / ms / 1e3:9.1f} MDoF/s  bitwise equal to the first: {same}", flush=True)
dev.close()
