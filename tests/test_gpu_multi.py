"""-m gpu, needs P >= 2 GPUs: every rank assembles / multiplies / solves on its RCB partition (ghost-layer assembly, halo
exchange, fused NVLink all-reduce) and is compared with the 1-rank CPU oracle in the same global numbering.  The ranks run
scripts/mgpu_check.py under torch.distributed.run (one process per GPU) and write their checks as JSON; the asserts are
here, one test per category and world size, so a driver with enough GPUs reports per-check results."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

_RESULTS = {}


def _run(world, tmp_path_factory):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run: gpurun --gpus {world} -- python -m pytest tests/test_gpu_multi.py -m gpu)")
    if world not in _RESULTS:
        out_dir = str(tmp_path_factory.mktemp(f"mgpu{world}"))
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
               "127.0.0.1", "--master-port", str(29533 + world), os.path.join(ROOT, "scripts", "mgpu_check.py")]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, MGPU_RESULT_DIR=out_dir))
        ranks = []
        for r in range(world):
            path = os.path.join(out_dir, f"rank{r}.json")
            assert os.path.exists(path), f"rank {r} wrote no result\n--- stdout tail ---\n{out.stdout[-2500:]}\n--- stderr tail ---\n{out.stderr[-2500:]}"
            ranks.append(json.load(open(path)))
        _RESULTS[world] = (out, ranks)
    return _RESULTS[world]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("category", ["assembly", "spmv", "gmres", "precond"])
def test_multi_rank_parity(world, category, tmp_path_factory):
    out, ranks = _run(world, tmp_path_factory)
    seen = 0
    for rk in ranks:
        assert rk["world"] == world
        for c in rk["checks"]:
            if c["category"] != category:
                continue
            seen += 1
            assert c["ok"], f"rank {rk['rank']} of {world}: {c}"
            if c["err"] is not None:
                assert c["err"] <= c["tol"], (rk["rank"], c)
    assert seen >= world, f"no {category} checks ran"
    if category == "spmv":      # the default kernel (variant 7) is among the halo SpMV checks, on every rank
        assert all(any(c["what"].endswith("variant 7") for c in rk["checks"]) for rk in ranks)
    if world == 4 and category == "assembly":
        assert max(rk["neighbors"] for rk in ranks) >= 2      # RCB corners: more than one neighbour
    assert "MGPU_CHECK PASS" in out.stdout


def test_two_rank_shim_twice_in_a_row(pkg):
    """The C++ shim (host/ns_app, the reference's main.cpp flow) on 2 ranks, started twice with the same MASTER_PORT: the NCCL id
    travels through a rendezvous file that belongs to one launch (nonce of the launcher, O_EXCL, removed when the communicator is
    up), so the second run cannot pick up the first run's id and hang; both runs give the 1-rank record."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import glob
    from conftest import mesh_path
    app = os.path.join(ROOT, "navier-stokes-dealii_b200", "host", "ns_app")
    env = dict(os.environ, NS_MESH=mesh_path("square_h0.1.msh"), NS_T="0.05", NS_NEUMANN_ID="1", NS_INLET_ID="0", NS_WALL_IDS="2,3")
    one = subprocess.run([app, "--history"], capture_output=True, text=True, env=env, timeout=300)
    assert one.returncode == 0, one.stderr[-1500:]
    ref = [json.loads(l) for l in one.stdout.splitlines() if l.startswith("{")]
    for attempt in range(2):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
               "127.0.0.1", "--master-port", "29577", app, "--history"]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, (attempt, r.stdout[-1500:], r.stderr[-1500:])
        recs = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        assert len(recs) == len(ref) >= 1
        for i, (a, b) in enumerate(zip(recs, ref)):
            assert (a["time_step"], a["newton"]) == (b["time_step"], b["newton"])
            # the first residual is a pure assembly result (rounding only); later ones follow GMRES solves stopped at 1e-2
            assert abs(a["residual"] - b["residual"]) <= (1e-9 if i == 0 else 0.2) * max(b["residual"], 1e-12), (i, a, b)
        assert not glob.glob("/tmp/ns_nccl_id.29577*"), "the rendezvous file must be removed once the communicator is up"
