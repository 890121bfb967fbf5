"""BASELINE.json configs[1] at its stated size: mesh-square-h0.012500.msh (12 800 cells, 58 403 DoFs), Stokes-initialised steady
Navier-Stokes, block preconditioners (hpp:520-639). usage: config2_run.py gpu|oracle OUT.json"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navier-stokes-dealii_b200")
backend, out = sys.argv[1], sys.argv[2]
prm = pkg.Parameters(mesh_path=os.path.join(ROOT, "tests", "golden", "square_h0.0125.msh"), nu=0.05, H=1.0, inlet_time_mode="constant",
                     neumann_id=1, inlet_id=0, wall_ids=(2, 3), clear_inlet_before_walls=True, use_mass=False,
                     preconditioner="block_diagonal", p_out=0.0, increment_bc="consistent", newton_max_iters=8)
mesh = pkg.Mesh.read_msh(prm.mesh_path, prm.surface_entity)
s = pkg.NavierStokesSolver(2, 1, 1.0, 1.0, prm, verbose=False)
if backend == "gpu":
    s.setup(mesh)
else:
    from oracle.oracle import Oracle
    s.mesh, s.dofs = mesh, pkg.Dofs(mesh)
    s.part = pkg.Part(s.dofs, 0)
    s.dev = Oracle(s.part)
    s._push_params(stokes=False)
t0 = time.perf_counter()
s.solve(stokes_init=True)
wall = time.perf_counter() - t0
sol = s.dev.get_solution()
print(f"{backend}: config 2 on square_h0.0125.msh: {s.dofs.n} DoFs, history {s.history}, {wall:.1f} s", flush=True)
json.dump({"backend": backend, "dofs": int(s.dofs.n), "wall_s": wall,
           "history": [[int(a), int(b), float(c), None if d is None else int(d)] for a, b, c, d in s.history]}, open(out, "w"))
np.save(out + ".sol.npy", sol)
