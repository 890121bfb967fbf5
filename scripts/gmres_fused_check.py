"""Fused (single cooperative kernel) GMRES vs the multi-kernel path on the meshes the reference ships."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG_S", "150")), exit=True)
pkg = importlib.import_module("navier-stokes-dealii_b200")
from conftest import analytic_state, mesh_path
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 0
m = pkg.Mesh.read_msh(mesh_path("cylinder_cmy.msh"))
if levels:
    m = m.refine(levels)
d = pkg.Dofs(m); part = pkg.Part(d, 0)
gd, gv = d.dirichlet_values([{11: True}, {11: True, 12: False, 13: False}], dict(u_m=1.5, H=0.41, time_factor=1.0))
dev = pkg.DeviceProblem(part, 0)
dev.set_params()
sol = analytic_state(d, 0.02)
dev.set_solution(sol); dev.set_solution_old(0.9 * sol)
dev.assemble(); dev.apply_dirichlet(gd, gv)
x0 = dev.get_delta()
print("cells", m.n_cells, "N", d.n, flush=True)
res = {}
for mode, tol, cap in ((0, 1e-2, 100000), (2, 1e-2, 100000), (0, 1e-8, 3000), (2, 1e-8, 3000)):
    dev.set_tuning(5, mode)
    for rep in range(2):
        dev.set_delta(x0)
        t = time.perf_counter()
        r = dev.solve(0, tol, cap, 30, 0, check=False)
        dt = time.perf_counter() - t
    h, x = dev.gmres_history(), dev.get_delta()
    res[(mode, tol)] = (r, h, x)
    print(f"fused={mode} tol={tol:g}: {r}  {dt:.4f} s  {1e6 * dt / max(r[0], 1):.1f} us/step  device ms {dev.phase_ms()['solve']:.2f}", flush=True)
for tol in (1e-2, 1e-8):
    (r0, h0, x0_), (r2, h2, x2) = res[(0, tol)], res[(2, tol)]
    k = min(len(h0), len(h2), 28)
    print(f"tol={tol:g}: steps {r0[0]} vs {r2[0]}; history first cycle rel diff {np.abs(h2[:k] / h0[:k] - 1).max():.2e}; "
          f"all common steps {np.abs(h2[:min(len(h0), len(h2))] / h0[:min(len(h0), len(h2))] - 1).max():.2e}; "
          f"delta rel diff {np.abs(x2 - x0_).max() / np.abs(x0_).max():.2e}", flush=True)
dev.close()
