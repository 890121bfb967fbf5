"""Quick device-vs-oracle check on one mesh (used during bring-up; the real tests are tests/ -m gpu)."""
import importlib, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navier-stokes-dealii_b200")
from oracle.oracle import Oracle

mesh_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests/golden/cylinder_cmy.msh")
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = pkg.Mesh.read_msh(mesh_path)
if levels: m = m.refine(levels)
d = pkg.Dofs(m); p = pkg.Part(d, 0)
print("cells", m.n_cells, "N", d.n, "nnz", p.nnz_jac, flush=True)
dev = pkg.DeviceProblem(p, 0)
o = Oracle(p); o.set_params()
xy = d.support_points()
sol = np.zeros(d.n)
sol[:d.n_u:2] = np.sin(np.pi*xy[:d.n_u:2,0])*np.cos(np.pi*xy[:d.n_u:2,1])
sol[1:d.n_u:2] = -np.cos(np.pi*xy[1:d.n_u:2,0])*np.sin(np.pi*xy[1:d.n_u:2,1])
sol[d.n_u:] = xy[d.n_u:,0]*xy[d.n_u:,1]
sol *= 0.02
old = 0.9*sol
for obj in (dev, o):
    obj.set_solution(sol); obj.set_solution_old(old); obj.assemble()
def rel(a, b):
    return np.abs(a-b).max()/max(np.abs(b).max(), 1e-300)
print("J   rel", rel(dev.get_matrix_values(), o.get_matrix_values()))
print("Mp  rel", rel(dev.get_pm_values(), o.get_pm_values()))
print("R   rel", rel(dev.get_residual(), o.get_residual()))
gd, gv = d.dirichlet_values([{11: True}, {11: True, 12: False, 13: False}], dict(u_m=1.5, H=0.41, time_factor=1.0))
ld, lv = p.localize_dirichlet(gd, gv)
dev.apply_dirichlet(ld, lv); o.apply_dirichlet(gd, gv)
print("J/bc rel", rel(dev.get_matrix_values(), o.get_matrix_values()))
print("R/bc rel", rel(dev.get_residual(), o.get_residual()))
print("norm", dev.residual_norm(), o.residual_norm())
x = np.random.default_rng(0).standard_normal(d.n)
print("spmv rel", rel(dev.spmv(x), o.spmv(x)))
x0 = dev.get_delta()
for g in (0, 1, 1):
    dev.set_tuning(2, g); dev.set_delta(x0)
x0 = dev.get_delta()
for g in (0, 1, 1):
    dev.set_tuning(2, g); dev.set_delta(x0)
    t = time.time(); r1 = dev.solve(0, 1e-2, 100000, 30, 0); t1 = time.time()-t
    print("graphs", g, "gmres", r1, "s", t1, "us/it", 1e6*t1/max(r1[0],1))
    print("graphs", g, "gmres", r1, "s", t1, "us/it", 1e6*t1/max(r1[0],1))
t = time.time(); r2 = o.solve(0, 1e-2, 100000, 30, 0); t2 = time.time()-t
print("gmres dev", r1, t1, "oracle", r2, t2)
print("delta rel", rel(dev.get_delta(), o.get_delta()))
h1, h2 = dev.gmres_history(), o.gmres_history()
n = min(len(h1), len(h2)); print("hist rel", np.abs(h1[:n]/h2[:n]-1).max(), len(h1), len(h2))
print("phase", dev.phase_ms(), dev.counters())
for w, name in [(0, "assembly"), (1, "spmv"), (2, "add_and_dot"), (3, "dot")]:
    print(name, "ms", dev.time_kernel(w, 20))
