"""Multi-GPU parity: run under  python -m torch.distributed.run --nproc-per-node P scripts/mgpu_check.py
Every rank owns one partition (RCB) with its ghost layer; results are compared with the 1-rank CPU
oracle on the same mesh in the SAME (part-major) global numbering."""
import importlib
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pkg = importlib.import_module("navier-stokes-dealii_b200")
from conftest import analytic_state, mesh_path  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def run_case(rank, world, local, uid, name, levels, calls, inlet, params, scale, preconds, state):
    m = pkg.Mesh.read_msh(mesh_path(name))
    if levels:
        m = m.refine(levels)
    cp = m.partition_rcb(world)
    d = pkg.Dofs(m, world, cp)
    part = pkg.Part(d, rank)
    dev = pkg.DeviceProblem(part, local)
    dev.comm_init(rank, world, uid)
    if dev.enable_peer_allreduce(dist) and rank == 0:
        print(f"{name}: Krylov all-reduces fused into the reduction kernels (NVLink peer mailboxes)", flush=True)
    own = part.l2g[: part.n_own]
    gd, gv = d.dirichlet_values(calls, dict(time_factor=1.0, **inlet))
    ld, lv = part.localize_dirichlet(gd, gv)

    class G:  # the whole problem in the same (part-major) numbering, for the oracle
        pass
    g = G()
    rp, col = d.sparsity(0)
    prp, pcol = d.sparsity(2)
    g.n_own_u, g.n_own_p, g.n_ghost_u, g.n_ghost_p = d.n_u, d.n_p, 0, 0
    g.n_own = d.n
    g.jac_rowptr, g.jac_col, g.pm_rowptr, g.pm_col = rp, col, prp, pcol
    g.nnz_jac, g.nnz_pm = len(col), len(pcol)
    g.n_cells, g.n_vertices = m.n_cells, m.n_vertices
    g.xy, g.cell_vertices, g.cell_dofs = m.xy.reshape(-1).copy(), m.cells.reshape(-1).copy(), d.cell_dofs.reshape(-1).copy()
    g.bface_cell, g.bface_face, g.bface_tag = m.boundary_faces()
    o = Oracle(g)
    # per-rank quantities of the reference (the Dirichlet diagonal d of apply_boundary_values, the ILU(0) blocks) are
    # taken rank by rank in the oracle too: virtual ranks = the owned ranges of the partition
    u_off = np.concatenate([[0], np.cumsum(d.part_n_u)])
    p_off = np.concatenate([[0], np.cumsum(d.part_n_p)])
    o.set_block_jacobi(u_off, p_off)

    sol = analytic_state(d, scale)
    for obj, s in ((dev, sol[own]), (o, sol)):
        obj.set_params(**params)
        obj.set_solution(s)
        obj.set_solution_old(0.9 * s)
        obj.assemble()

    def check(what, a, b, tol, category="assembly"):
        err = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
        good = err <= tol
        state["ok"] &= good
        state["checks"].append({"mesh": name, "category": category, "what": what, "err": err, "tol": tol, "ok": bool(good)})
        print(f"[rank {rank}] {time.strftime('%H:%M:%S')} {name}: {what:28s} rel err {err:.3e} {'ok' if good else 'FAIL'}", flush=True)

    def record(what, good, category, detail):
        state["ok"] &= bool(good)
        state["checks"].append({"mesh": name, "category": category, "what": what, "err": None, "tol": None, "ok": bool(good),
                                "detail": detail})
        if not good:
            print(f"[rank {rank}] {name}: {what} FAIL {detail}", flush=True)

    Jo = o.get_matrix_values()
    jo_rows = np.concatenate([_reorder(part, Jo, rp, col, gr) for gr in own])
    check("J (owned rows)", dev.get_matrix_values(), jo_rows, 1e-12)
    check("R", dev.get_residual(), o.get_residual()[own], 1e-12)
    dev.apply_dirichlet(ld, lv)
    o.apply_dirichlet(gd, gv)
    check("R after Dirichlet", dev.get_residual(), o.get_residual()[own], 1e-12)
    check("||R|| (allreduce)", np.array([dev.residual_norm()]), np.array([o.residual_norm()]), 1e-12)
    x = np.random.default_rng(5).standard_normal(d.n)
    for variant in (0, 1, 4, 7):       # 7 = the default
        dev.set_tuning(0, variant)
        check(f"SpMV with halo, variant {variant}", dev.spmv(x[own]), o.spmv(x)[own], 1e-12, "spmv")
    state["neighbors"] = max(state.get("neighbors", 0), int(part.n_neighbors))
    rd = dev.solve(0, 1e-2, 100000, 30, 0, check=False)
    ro = o.solve(0, 1e-2, 100000, 30, 0)
    print(f"[rank {rank}] {name}: GMRES identity: device {rd} oracle {ro}", flush=True)
    record("GMRES identity step counts", rd[2] == ro[2] and abs(rd[0] - ro[0]) <= max(2, 0.1 * ro[0]), "gmres", f"device {rd} oracle {ro}")
    info = dev.last_solve_info()
    record("GMRES identity ran the multi-kernel path with SpMV 7", (not info["fused"]) and info["spmv_variant"] == 7, "gmres", str(info))
    h1, h2 = dev.gmres_history(), o.gmres_history()
    k = min(28, len(h1), len(h2))
    check("GMRES history (first cycle)", h1[:k], h2[:k], 1e-9, "gmres")
    if rd[0] == ro[0] and rd[0] < 400:
        check("delta", dev.get_delta(), o.get_delta()[own], 1e-6, "gmres")
    # block preconditioners: per-rank ILU(0) == block-Jacobi ILU(0) in the oracle (virtual ranks set above)
    for precond in preconds:
        dev.set_delta(np.zeros(part.n_own))
        o.set_delta(np.zeros(d.n))
        rd = dev.solve(precond, 1e-6, 2000, 30, 0, check=False)
        ro = o.solve(precond, 1e-6, 2000, 30, 0)
        print(f"[rank {rank}] {name}: GMRES precond {precond}: device {rd} oracle {ro}", flush=True)
        record(f"GMRES precond {precond} step counts", rd[2] == ro[2] == 0 and rd[0] == ro[0], "precond", f"device {rd} oracle {ro}")
        check(f"delta precond {precond}", dev.get_delta(), o.get_delta()[own], 1e-6, "precond")
    dev.close()


def main():
    # a stuck collective must not hang the GPU box: dump every thread's Python stack and exit
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("MGPU_WATCHDOG_S", "240")), exit=True)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    state = {"ok": True, "checks": []}
    for case in (("cylinder_cmy.msh", 0, [{11: True}, {11: True, 12: False, 13: False}], dict(u_m=1.5, H=0.41), dict(), 0.02, ()),
                 ("square_h0.05.msh", 0, [{0: True}, {2: False, 3: False}], dict(u_m=1.5, H=1.0), dict(nu=0.01, neumann_id=1), 0.05,
                  (2, 1))):
        uid = [pkg.DeviceProblem.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        run_case(rank, world, local, uid[0], case[0], case[1], case[2], case[3], case[4], case[5], case[6], state)
    t = torch.tensor([1.0 if state["ok"] else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out_dir = os.environ.get("MGPU_RESULT_DIR")
    if out_dir:      # per-rank machine-readable results for tests/test_gpu_multi.py
        import json
        with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
            json.dump({"rank": rank, "world": world, "neighbors": state.get("neighbors", 0), "checks": state["checks"]}, f)
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_CHECK", "PASS" if t.item() == 1.0 else "FAIL", flush=True)
    sys.exit(0 if t.item() == 1.0 else 1)


def _reorder(part, Jo, rp, col, gr):
    """values of global row gr in the LOCAL column order of the rank (ascending local ids)."""
    gcols = col[rp[gr]:rp[gr + 1]]
    lcols = _loc(part, gcols)
    return Jo[rp[gr]:rp[gr + 1]][np.argsort(lcols, kind="stable")]


def _loc(part, gcols):
    # cached ON the part object (a cache keyed by id(part) handed a later case the map of a collected object)
    g2l = getattr(part, "_g2l_map", None)
    if g2l is None:
        g2l = {int(g): i for i, g in enumerate(part.l2g)}
        part._g2l_map = g2l
    return np.array([g2l[int(c)] for c in gcols])


if __name__ == "__main__":
    main()
