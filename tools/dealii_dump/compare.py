"""Check this repository against a dump written by tools/dealii_dump (the reference run on a deal.II machine).

    python tools/dealii_dump/compare.py DUMP_DIR [--mesh tests/golden/cylinder_cmy.msh] [--gpu]

Checks, in the order of north_star's parity bar:
  1. dof maps: the 15 global dof ids of every cell              bit-exact
  2. sparsity patterns of the Jacobian and the pressure mass     bit-exact (as sets per row, and column order)
  3. assembled Jacobian / pressure mass entries, residual        1e-12 relative to the row / vector maximum
  4. Newton increment delta of the first solve_system()          1e-8 relative (identity-preconditioned GMRES)
The CPU oracle is checked always; `--gpu` checks the CUDA path through the C-ABI as well.
`--self-test DIR` writes a dump in the same format FROM THE ORACLE (format check of this script only)."""
import argparse
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("navier-stokes-dealii_b200")
from oracle.oracle import Oracle  # noqa: E402

# the reference's hard-wired set-up (cpp:351-373: ids 11 inlet, 12/13 walls; hpp:473-474 inlet data)
CALLS = [{11: True}, {11: True, 12: False, 13: False}]
INLET = dict(u_m=1.5, H=0.41, time_factor=0.0)   # the reference never calls set_time: sin(pi*0/8) = 0


def load_cells(path):
    a = np.loadtxt(path)
    return a[:, :6], a[:, 6:].astype(np.int64)


def load_pattern(path):
    rows = {}
    for line in open(path):
        r, _, cols = line.partition(":")
        rows[int(r)] = np.array(cols.split(), dtype=np.int64)
    return rows


def load_triplets(path):
    a = np.loadtxt(path, ndmin=2)
    return a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2]


def load_vector(path, n):
    a = np.loadtxt(path, ndmin=2)
    v = np.zeros(n)
    v[a[:, 0].astype(np.int64)] = a[:, 1]
    return v


def build(mesh_path):
    m = pkg.Mesh.read_msh(mesh_path)
    d = pkg.Dofs(m)
    part = pkg.Part(d, 0)
    gd, gv = d.dirichlet_values(CALLS, INLET)
    return m, d, part, gd, gv


def run_backend(obj, d, gd, gv):
    obj.set_params()
    obj.set_solution(np.zeros(d.n))       # FunctionU0 = 0 (hpp:438-457)
    obj.push_time_level()
    obj.assemble()
    obj.apply_dirichlet(gd, gv)
    J, Mp, R = obj.get_matrix_values(), obj.get_pm_values(), obj.get_residual()
    obj.solve(0, 1e-2, 100000, 30, 0)
    return J, Mp, R, obj.get_delta()


def self_test(out, mesh_path):
    m, d, part, gd, gv = build(mesh_path)
    os.makedirs(out, exist_ok=True)
    xy = m.xy.reshape(-1, 2)[m.cells.reshape(-1, 3)].reshape(-1, 6)
    np.savetxt(os.path.join(out, "cells.txt"), np.hstack([xy, d.cell_dofs.reshape(-1, 15)]), fmt=["%.17g"] * 6 + ["%d"] * 15)
    J, Mp, R, delta = run_backend(Oracle(part), d, gd, gv)
    for kind, vals, pat, valf in ((0, J, "pattern.txt", "jacobian.txt"), (2, Mp, "pm_pattern.txt", "pm.txt")):
        rp, col = d.sparsity(kind)
        with open(os.path.join(out, pat), "w") as fp, open(os.path.join(out, valf), "w") as fv:
            for r in range(d.n):
                fp.write(f"{r}:" + "".join(f" {c}" for c in col[rp[r]:rp[r + 1]]) + "\n")
                for p in range(rp[r], rp[r + 1]):
                    fv.write(f"{r} {col[p]} {vals[p]:.17g}\n")
    for name, v in (("residual.txt", R), ("delta.txt", delta)):
        np.savetxt(os.path.join(out, name), np.column_stack([np.arange(d.n), v]), fmt=["%d", "%.17g"])
    print("self-test dump written to", out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump")
    ap.add_argument("--mesh", default=os.path.join(ROOT, "tests", "golden", "cylinder_cmy.msh"))
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--self-test", action="store_true")
    a = ap.parse_args()
    if a.self_test:
        return self_test(a.dump, a.mesh)
    m, d, part, gd, gv = build(a.mesh)
    ok = True

    def report(what, good, detail=""):
        nonlocal ok
        ok &= bool(good)
        print(f"{what:58s} {'ok' if good else 'FAIL'} {detail}")

    # 1. dof maps (cells in the same order: deal.II iterates the cells in the order read_msh created them)
    xy, ids = load_cells(os.path.join(a.dump, "cells.txt"))
    mine_xy = m.xy.reshape(-1, 2)[m.cells.reshape(-1, 3)].reshape(-1, 6)
    report("cell order and vertex coordinates", xy.shape == mine_xy.shape and np.abs(xy - mine_xy).max() < 1e-14)
    report("cell -> 15 global dof ids (bit-exact)", ids.shape == (m.n_cells, 15) and np.array_equal(ids, d.cell_dofs.reshape(-1, 15)))
    # 2. patterns
    for kind, pat in ((0, "pattern.txt"), (2, "pm_pattern.txt")):
        rp, col = d.sparsity(kind)
        rows = load_pattern(os.path.join(a.dump, pat))
        same_set = all(np.array_equal(np.sort(rows.get(r, np.zeros(0, np.int64))), col[rp[r]:rp[r + 1]]) for r in range(d.n))
        same_order = all(np.array_equal(rows.get(r, np.zeros(0, np.int64)), col[rp[r]:rp[r + 1]]) for r in range(d.n))
        report(f"{pat}: same entries per row (bit-exact)", same_set)
        report(f"{pat}: same stored column order", same_order, "" if same_order else "(Epetra local-index order differs: informational)")
    # 3./4. values
    backends = [("oracle", Oracle(part))]
    if a.gpu:
        backends.append(("cuda", pkg.DeviceProblem(part, 0)))
    for name, obj in backends:
        J, Mp, R, delta = run_backend(obj, d, gd, gv)
        for kind, mine, valf in ((0, J, "jacobian.txt"), (2, Mp, "pm.txt")):
            rp, col = d.sparsity(kind)
            r, c, v = load_triplets(os.path.join(a.dump, valf))
            pos = {}
            for row in np.unique(r):
                pos[row] = {int(cc): int(p) for p, cc in zip(range(rp[row], rp[row + 1]), col[rp[row]:rp[row + 1]])}
            ref = np.zeros_like(mine)
            for rr, cc, vv in zip(r, c, v):
                ref[pos[rr][int(cc)]] = vv
            scale = np.maximum.reduceat(np.abs(ref), rp[:-1][np.diff(rp) > 0]).max() if len(ref) else 1.0
            err = np.abs(mine - ref).max() / max(scale, 1e-300)
            report(f"{name}: {valf} entries", err <= 1e-12, f"rel err {err:.2e}")
        ref_R = load_vector(os.path.join(a.dump, "residual.txt"), d.n)
        err = np.abs(R - ref_R).max() / max(np.abs(ref_R).max(), 1e-300)
        report(f"{name}: residual", err <= 1e-12, f"rel err {err:.2e}")
        ref_d = load_vector(os.path.join(a.dump, "delta.txt"), d.n)
        err = np.abs(delta - ref_d).max() / max(np.abs(ref_d).max(), 1e-300)
        report(f"{name}: Newton increment after solve_system", err <= 1e-8, f"rel err {err:.2e}")
    print("DEALII_COMPARE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
