"""world_size-2 gloo test (CPU): the host side of the multi-GPU path — partition, part-major DoF
numbering, ghost layer, halo plan — exercised with real point-to-point messages, and the rank-local
CSR rows checked against the global operator through a distributed SpMV in numpy."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, mesh_path


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib
    pkg = importlib.import_module("navier-stokes-dealii_b200")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = pkg.Mesh.read_msh(mesh_path("square_h0.05.msh"))
        cp = m.partition_rcb(world)
        d = pkg.Dofs(m, world, cp)
        part = pkg.Part(d, rank)
        # a global vector every rank can evaluate: x_g = f(global id)
        f = lambda g: np.sin(0.37 * g) + 1e-3 * g
        x = np.zeros(part.n_loc)
        x[: part.n_own] = f(part.l2g[: part.n_own])
        # halo exchange following the plan (what K8 does with ncclSend/ncclRecv)
        reqs, bufs = [], []
        for k, nb in enumerate(part.neighbors):
            s = torch.from_numpy(x[part.send_idx[part.send_ptr[k]:part.send_ptr[k + 1]]].copy())
            r = torch.zeros(int(part.recv_ptr[k + 1] - part.recv_ptr[k]), dtype=torch.float64)
            reqs += [dist.isend(s, int(nb)), dist.irecv(r, int(nb))]
            bufs.append((k, r, s))
        for rq in reqs:
            rq.wait()
        for k, r, _ in bufs:
            x[part.recv_idx[part.recv_ptr[k]:part.recv_ptr[k + 1]]] = r.numpy()
        ok_halo = np.array_equal(x, f(part.l2g))
        # distributed SpMV with the pattern (values = 1/(1+|i-j| mod 7)) vs the global pattern rows
        rp, col = d.sparsity(0)
        ok_rows = True
        y_loc = np.zeros(part.n_own)
        for i in range(part.n_own):
            lc = part.jac_col[part.jac_rowptr[i]:part.jac_rowptr[i + 1]]
            gi = part.l2g[i]
            gc = part.l2g[lc]
            w = 1.0 / (1 + (np.abs(gi - gc) % 7))
            y_loc[i] = np.dot(w, x[lc])
            if i % 97 == 0:
                ok_rows &= sorted(gc.tolist()) == col[rp[gi]:rp[gi + 1]].tolist()
        # global reference on every rank
        xg = f(np.arange(d.n))
        y_ref = np.array([np.dot(1.0 / (1 + (np.abs(g - col[rp[g]:rp[g + 1]]) % 7)), xg[col[rp[g]:rp[g + 1]]])
                          for g in part.l2g[: part.n_own]])
        ok_spmv = np.allclose(y_loc, y_ref, rtol=1e-13, atol=0)
        # allreduce of a partial dot product = the global one (K5 reductions)
        t = torch.tensor([float(np.dot(x[: part.n_own], x[: part.n_own]))], dtype=torch.float64)
        dist.all_reduce(t)
        ok_dot = abs(t.item() - float(np.dot(xg, xg))) <= 1e-12 * float(np.dot(xg, xg))
        q.put((rank, bool(ok_halo), bool(ok_rows), bool(ok_spmv), bool(ok_dot), part.n_neighbors))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_halo_and_spmv_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert len(res) == 2
    for rank, ok_halo, ok_rows, ok_spmv, ok_dot, nn in res:
        assert ok_halo and ok_rows and ok_spmv and ok_dot and nn == 1, res


class _FakeLib:
    def __init__(self, log):
        self.log = log

    def nsg_comm_release_peers(self, h):
        self.log.append("release")
        return 0


def _peer_worker(rank, world, port, q, failing_rank):
    """Host logic of DeviceProblem.enable_peer_allreduce on a stub (no GPU): the switch to the fused all-reduce is a
    COLLECTIVE decision - if any rank cannot export/map a mailbox, every rank stays on NCCL and releases its mappings."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import importlib
    device = importlib.import_module("navier-stokes-dealii_b200.device")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        log = []

        class Stub:
            _L, _h = _FakeLib(log), object()

            def comm_ipc_handle(self):
                return bytes([rank]) * 64

            def comm_set_peers(self, handles):
                log.append(("set", [h[0] for h in handles]))
                if rank == failing_rank:
                    raise RuntimeError("cudaIpcOpenMemHandle: peer access is not supported between these two devices")

        stub = Stub()
        os.environ.pop("NSG_NO_PEER_AR", None)
        got = device.DeviceProblem.enable_peer_allreduce(stub, dist)
        q.put((rank, got, log, getattr(stub, "_peer_dist", None) is not None))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("failing_rank", [-1, 1])
def test_peer_allreduce_switch_is_collective_gloo(failing_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() + 7 * (failing_rank + 2)) % 90
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q, failing_rank)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
    for rank, got, log, has_dist in res:
        assert ("set", [0, 1]) in log                          # handles arrive in rank order on every rank
        if failing_rank < 0:
            assert got is True and has_dist and "release" not in log
        else:
            assert got is False and not has_dist and "release" in log, res
