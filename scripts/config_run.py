"""BASELINE.json configs 1 and 3 through the Python mirror of the reference class (solver.py: same setup()/solve() call order
as src/main.cpp), on the CUDA path or on the CPU oracle plugged into the same driver; writes the Newton/GMRES history and
the drag/lift history as JSON.
  python scripts/config_run.py --backend gpu|oracle --config 1|3 --steps N [--levels L] --out FILE
config 1: flow past the cylinder of mesh2d.msh (surface entity 5), Re = 20, dt = 0.05, reference solver settings
          (GMRES(28) identity, 1e-2 relative; Newton 1e-2 absolute; cpp:566,593-594), 10 steps.
config 3: the same domain at Re = 100 (nu = 0.01, mean inflow 1, D = 1), long run with drag/lift after every step."""
import argparse, importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navier-stokes-dealii_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--backend", default="gpu", choices=["gpu", "oracle"])
ap.add_argument("--config", type=int, default=1)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--levels", type=int, default=0)
ap.add_argument("--dt", type=float, default=0.05)
ap.add_argument("--rel-tol", type=float, default=1e-2)
ap.add_argument("--newton-max", type=int, default=8)
ap.add_argument("--precond", default="identity")
ap.add_argument("--out", default="")
a = ap.parse_args()
nu = {1: 0.05, 3: 0.01}[a.config]
prm = pkg.Parameters(mesh_path=os.path.join(ROOT, "tests", "golden", "cylinder_mesh2d.msh"), surface_entity=5, nu=nu, u_m=1.5, H=4.1,
                     inlet_y0=-2.0, inlet_time_mode="constant", preconditioner=a.precond, force_boundary_id=3, p_out=0.0,
                     increment_bc="consistent", neumann_id=1, inlet_id=0, wall_ids=(2, 3), gmres_rel_tol=a.rel_tol,
                     gmres_max_iters=1000000, refine_levels=a.levels, newton_max_iters=a.newton_max)
mesh = pkg.Mesh.read_msh(prm.mesh_path, prm.surface_entity)
mesh.tag_boundary_box(0, 1, 2, 3)
if a.levels:
    mesh = mesh.refine(a.levels)
s = pkg.NavierStokesSolver(2, 1, a.steps * a.dt, a.dt, prm, verbose=False)
if a.backend == "gpu":
    s.setup(mesh)
else:
    from oracle.oracle import Oracle
    s.mesh = mesh
    s.dofs = pkg.Dofs(mesh)
    s.part = pkg.Part(s.dofs, 0)
    s.dev = Oracle(s.part)
    s._push_params(stokes=False)
t0 = time.perf_counter()
s.solve()
wall = time.perf_counter() - t0
out = {"backend": a.backend, "config": a.config, "steps": a.steps, "levels": a.levels, "dt": a.dt, "nu": nu, "rel_tol": a.rel_tol,
       "cells": int(mesh.n_cells), "dofs": int(s.dofs.n), "wall_s": wall,
       "history": [[int(x[0]), int(x[1]), float(x[2]), None if x[3] is None else int(x[3])] for x in s.history],
       "forces": [[float(t), float(fx), float(fy)] for t, fx, fy in s.force_history],
       "solution_norm": float(np.linalg.norm(s.dev.get_solution()))}
its = sum(x[3] or 0 for x in s.history)
print(f"{a.backend}: config {a.config} L{a.levels} {mesh.n_cells} cells {s.dofs.n} DoFs, {a.steps} steps, {len(s.history)} Newton its, {its} GMRES steps, "
      f"{wall:.1f} s; last force {s.force_history[-1] if s.force_history else None}", flush=True)
if a.out:
    np.save(a.out + ".sol.npy", s.dev.get_solution())
    json.dump(out, open(a.out, "w"))
