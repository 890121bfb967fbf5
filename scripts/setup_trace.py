"""Where the one-time setup of the default bench workload goes (host topology stages + NSG_TRACE marks of libnsg)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NSG_TRACE"] = "1"
import numpy as np
import bench
pkg = importlib.import_module("navier-stokes-dealii_b200")
L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t = [time.perf_counter()]
def mark(what):
    t.append(time.perf_counter()); print(f"[setup trace] {what}: {t[-1] - t[-2]:.2f} s", flush=True)
fn, ent, geo, calls, neumann, inlet = bench.MESHES["mesh2d"]
m = pkg.Mesh.read_msh(os.path.join(ROOT, "tests", "golden", fn), ent); m.tag_boundary_box(0, 1, 2, 3); mark("read")
m = m.refine(L); mark(f"refine x{L} ({m.n_cells} cells)")
d = pkg.Dofs(m); mark("dofs")
part = pkg.Part(d, 0, patterns=False); mark("part (no patterns: built on the device)")
gd, gv = d.dirichlet_values(calls, dict(time_factor=1.0, **inlet)); ld, lv = part.localize_dirichlet(gd, gv); mark("dirichlet list")
xy = d.support_points(); sol = bench.analytic_state(xy, d.n_u)[part.l2g[: part.n_own]]; mark("support points + state")
dev = pkg.DeviceProblem(part, 0); mark("DeviceProblem (set_pattern + set_mesh)")
dev.set_solution(sol); dev.set_solution_old(0.9 * sol); mark("set_solution x2")
print(f"[setup trace] total {t[-1] - t[0]:.2f} s", flush=True)
dev.close()
