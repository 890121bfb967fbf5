"""Converged preconditioned Newton-step solve (hpp:520-639) on the largest mesh where the reference's block preconditioners
converge: the steady Navier-Stokes system of BASELINE.json configs[1] (square mesh, nu = 0.05) at `levels` refinements of
mesh-square-h0.012500.msh.  Prints one JSON object: outer/inner iteration counts, time per solve, time and bytes per ILU(0)
apply (the two triangular solves), time of the ILU factorisation.
usage: precond_bench.py [levels] [precond: 1 block-diagonal | 2 block-triangular]"""
import importlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("navier-stokes-dealii_b200")
from conftest import analytic_state, mesh_path
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 1
preconds = [int(x) for x in sys.argv[2].split(",") if x not in ("", "-")] if len(sys.argv) > 2 else [2, 1]   # "-" = time the ILU applies only
m = pkg.Mesh.read_msh(mesh_path("square_h0.0125.msh"))
if levels:
    m = m.refine(levels)
d = pkg.Dofs(m); part = pkg.Part(d, 0)
gd, gv = d.dirichlet_values([{0: True}, {2: False, 3: False}], dict(time_factor=1.0, u_m=1.5, H=1.0))
dev = pkg.DeviceProblem(part, 0)
dev.set_params(nu=0.05, neumann_id=1, use_mass=0, p_out=0.0)
dev.set_solution(analytic_state(d, 0.05)); dev.set_solution_old(analytic_state(d, 0.045))
dev.assemble(); dev.apply_dirichlet(gd, gv)
out = {"mesh": "square_h0.0125.msh", "levels": levels, "cells": int(m.n_cells), "dofs": int(d.n), "n_u": int(d.n_u), "nnz_jac": int(part.nnz_jac)}
rp = part.jac_rowptr
nnz_A = int(sum(1 for _ in ()))  # placeholder, computed below
cols = part.jac_col
rows_u = d.n_u
# non-zeros of the velocity block A = J[0:n_u, 0:n_u]
nnz_A = int((cols[: rp[rows_u]] < d.n_u).sum())
x = np.random.default_rng(0).standard_normal(d.n_u)
dev.ilu_apply(0, x)      # factorise + warm up
t = time.perf_counter()
for _ in range(5):
    dev.ilu_apply(0, x)
t_apply = (time.perf_counter() - t) / 5
out["ilu_apply_A"] = {"ms": 1e3 * t_apply, "nnz": nnz_A, "algorithmic_bytes": 12 * nnz_A + 24 * d.n_u,
                      "GB_per_s": (12 * nnz_A + 24 * d.n_u) / t_apply / 1e9,
                      "note": "forward + backward triangular solve in natural (DoF) order as Ifpack does; one launch each, rows wait on the "
                              "completion stamps of the rows they depend on; bound by the length of the dependency chain, not by bytes"}
out["ilu_apply_A_device_ms"] = {}
for variant, label in ((0, "one launch per level"), (1, "stamped single launch, all SMs"), (2, "one CTA, shared-memory window"), (-1, "library's choice")):
    dev.set_tuning(4, variant)
    dev.time_kernel(6, 2)
    out["ilu_apply_A_device_ms"][label] = dev.time_kernel(6, 20)
out["ilu_apply_Mp_device_ms"] = {}
for variant, label in ((1, "stamped single launch, all SMs"), (2, "one CTA, shared-memory window")):
    dev.set_tuning(4, variant)
    dev.time_kernel(7, 2)
    out["ilu_apply_Mp_device_ms"][label] = dev.time_kernel(7, 20)
dev.set_tuning(4, -1)
for pc in preconds:
    dev.set_delta(np.zeros(d.n))
    t = time.perf_counter()
    its, res, rc = dev.solve(pc, 1e-6, 2000, 30, 0, check=False)
    out[{1: "block_diagonal", 2: "block_triangular"}[pc]] = {"outer_gmres_steps": int(its), "rc": int(rc), "last_residual": float(res),
                                                             "solve_s": time.perf_counter() - t, "device_ms": dev.phase_ms()["solve"]}
if not preconds:
    print(json.dumps(out), flush=True)
    dev.close()
    sys.exit(0)
dev.set_delta(np.zeros(d.n))
t = time.perf_counter()
its, res, rc = dev.solve(0, 1e-6, 20000, 30, 0, check=False)
out["identity"] = {"outer_gmres_steps": int(its), "rc": int(rc), "last_residual": float(res), "solve_s": time.perf_counter() - t}
print(json.dumps(out), flush=True)
dev.close()
