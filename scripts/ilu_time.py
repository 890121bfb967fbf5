"""Preconditioned solve timing: level-scheduled vs single-launch ILU(0) triangular solves."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("navier-stokes-dealii_b200")
from conftest import analytic_state, mesh_path
name, levels = (sys.argv[1] if len(sys.argv) > 1 else "square_h0.0125.msh"), int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = pkg.Mesh.read_msh(mesh_path(name))
if levels: m = m.refine(levels)
d = pkg.Dofs(m); part = pkg.Part(d, 0)
calls = [{0: True}, {2: False, 3: False}]
gd, gv = d.dirichlet_values(calls, dict(time_factor=1.0, u_m=1.5, H=1.0))
dev = pkg.DeviceProblem(part, 0)
dev.set_params(nu=0.01, neumann_id=1)
dev.set_solution(analytic_state(d, 0.05)); dev.set_solution_old(analytic_state(d, 0.045))
dev.assemble(); dev.apply_dirichlet(gd, gv)
print("cells", m.n_cells, "N", d.n, flush=True)
x = np.random.default_rng(0).standard_normal(d.n_u)
for v in [int(a) for a in (sys.argv[3] if len(sys.argv) > 3 else "0,1").split(",")]:
    dev.set_tuning(4, v)
    dev.ilu_apply(0, x)
    t = time.perf_counter()
    for _ in range(10): y = dev.ilu_apply(0, x)
    t_apply = (time.perf_counter() - t) / 10
    r, t_solve = None, 0.0
    if os.environ.get("ILU_TIME_SOLVE"):
        dev.set_delta(np.zeros(d.n))
        t = time.perf_counter(); r = dev.solve(2, 1e-6, 2000, 30, 0, check=False); t_solve = time.perf_counter() - t
    print(f"ilu variant {v}: apply(A) incl. copies {1e3 * t_apply:.3f} ms; block-triangular GMRES {r} in {t_solve:.3f} s", flush=True)
dev.close()
