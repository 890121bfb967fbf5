"""ncu target: a few launches of assembly variant 2 on a level-L mesh."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
av = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m, d, part, (ld, lv), neumann, sol = bench.build_problem(pkg, "cmy", L, 1, 0)
dev = pkg.DeviceProblem(part, 0)
dev.set_params(neumann_id=neumann)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
dev.set_tuning(1, av)
print("ms", dev.time_kernel(0, 3))
dev.close()
