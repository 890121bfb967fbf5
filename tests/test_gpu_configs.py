"""-m gpu: the BASELINE.json configurations that fit a unit test, driven through the Python mirror of the
reference class (same setup/solve call order as src/main.cpp) on BOTH back ends: the CUDA path and the
CPU oracle plugged into the same driver."""
import numpy as np
import pytest

from conftest import mesh_path
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu


def run_both(pkg, prm, T, dt, stokes_init=False, box_tags=None):
    mesh = pkg.Mesh.read_msh(prm.mesh_path, prm.surface_entity)
    if box_tags:
        mesh.tag_boundary_box(*box_tags)     # untagged mesh: geometric boundary ids
    s = pkg.NavierStokesSolver(2, 1, T, dt, prm, verbose=False)
    s.setup(mesh)
    o = pkg.NavierStokesSolver(2, 1, T, dt, prm, verbose=False)
    o.mesh, o.dofs = s.mesh, s.dofs
    o.part = pkg.Part(s.dofs, 0)           # with the host-built patterns (the device builds its own from the cells)
    o.dev = Oracle(o.part)                 # same driver, CPU oracle behind it

    def checked(solve):
        def f(*a, **k):
            r = solve(*a, **k)
            assert r[2] == 0, r
            return r
        return f
    o.dev.solve = checked(o.dev.solve)
    o._push_params(stokes=False)
    s.solve(stokes_init=stokes_init)
    o.solve(stokes_init=stokes_init)
    return s, o


def test_config2_stokes_initialised_steady_ns(pkg):
    """configs[1]: square mesh, Stokes solve as initial guess (cpp:636-644 with the Stokes ids 0/1/2/3), then
    steady Navier-Stokes Newton iterations (no mass term). Block-diagonal preconditioner (hpp:520-572) and
    increment BCs g - u^k, without which Newton cannot converge for g != 0 (see Parameters.increment_bc)."""
    prm = pkg.Parameters(mesh_path=mesh_path("square_h0.1.msh"), nu=0.05, H=1.0, inlet_time_mode="constant", neumann_id=1,
                         inlet_id=0, wall_ids=(2, 3), clear_inlet_before_walls=True, use_mass=False,
                         preconditioner="block_diagonal", p_out=0.0, increment_bc="consistent", newton_max_iters=8)
    s, o = run_both(pkg, prm, T=1.0, dt=1.0, stokes_init=True)
    assert [(a, b, d) for a, b, _, d in s.history] == [(a, b, d) for a, b, _, d in o.history]
    for (_, _, r1, _), (_, _, r2, _) in zip(s.history, o.history):
        assert abs(r1 - r2) <= 1e-8 * max(r2, 1e-2)
    xs, xo = s.dev.get_solution(), o.dev.get_solution()
    assert np.abs(xs - xo).max() <= 1e-8 * np.abs(xo).max()
    assert s.history[-1][2] <= 1e-2 and len(s.history) == 3          # Newton converged in 3 iterations
    xy = s.dofs.support_points()
    mid = np.isclose(xy[:s.dofs.n_u:2, 0], 0.5) & np.isclose(xy[:s.dofs.n_u:2, 1], 0.5)
    assert xs[:s.dofs.n_u:2][mid].min() > 0.5                        # developed channel flow


def test_config1_unsteady_cylinder_re20_ten_steps(pkg):
    """configs[0] at its stated length: flow past the cylinder (surface entity 5 of mesh2d.msh), Re = 20 (nu = 0.05, mean inflow 1,
    D = 1), inlet on, TEN implicit-Euler steps of Newton + GMRES(28, identity), drag/lift on the cylinder after every step.
    The linear solves run to 1e-12 instead of the reference's 1e-2 (cpp:566): with the inlet on, the reference's own settings do
    not give a usable Newton iteration - the oracle needs 4, 5, 7, 9, 10+ Newton iterations in time steps 1..5 and stalls above
    the 1e-2 Newton tolerance from step 5 on (profiles/r02_summary.md) - and a comparison at 1e-8 must not be limited by where
    two implementations stop along 50 000-step restarted solves (at a solver tolerance of 1e-10 that leaves 1.1e-8 in the
    iterate).  33 Newton iterations, 1.15 million GMRES steps on either side."""
    prm = pkg.Parameters(mesh_path=mesh_path("cylinder_mesh2d.msh"), surface_entity=5, nu=0.05, u_m=1.5, H=4.1, inlet_y0=-2.0,
                         inlet_time_mode="constant", preconditioner="identity", force_boundary_id=3, p_out=0.0,
                         increment_bc="consistent", neumann_id=1, inlet_id=0, wall_ids=(2, 3), newton_max_iters=8,
                         gmres_rel_tol=1e-12, gmres_max_iters=1000000)
    s, o = run_both(pkg, prm, T=0.5, dt=0.05, box_tags=(0, 1, 2, 3))
    assert len(s.force_history) == len(o.force_history) == 10
    assert [(a, b, d is None) for a, b, _, d in s.history] == [(a, b, d is None) for a, b, _, d in o.history]
    assert max(a for a, _, _, _ in s.history) == 10 and all(r <= 1e-2 for _, _, r, d in s.history if d is None)   # every step converged
    fs, fo = np.array(s.force_history), np.array(o.force_history)
    print("forces", fs[-1].tolist(), fo[-1].tolist())
    assert np.abs(fs[:, 1:] - fo[:, 1:]).max() <= 1e-8 * np.abs(fo[:, 1]).max()      # drag and lift history
    for (_, _, r1, _), (_, _, r2, _) in zip(s.history, o.history):
        assert abs(r1 - r2) <= 1e-8 * max(r2, 1.0)
    xs, xo = s.dev.get_solution(), o.dev.get_solution()
    assert np.abs(xs - xo).max() <= 1e-8 * np.abs(xo).max()
    assert s.dev.last_solve_info()["fused"]          # 1 451 unknowns: the cooperative single-kernel solver


def test_error_paths(pkg):
    import importlib
    lib = importlib.import_module("navier-stokes-dealii_b200._lib")
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    d = pkg.Dofs(m)
    part = pkg.Part(d, 0)
    dev = pkg.DeviceProblem(part, 0)
    with pytest.raises(lib.NsgError, match="not a locally owned row"):
        dev.apply_dirichlet(np.array([d.n + 5], np.int32), np.array([1.0]))
    dev.apply_dirichlet(np.zeros(0, np.int32), np.zeros(0))            # empty list is a no-op
    with pytest.raises(lib.NsgError, match="n_tmp_vectors"):
        dev.solve(0, 1e-2, 10, 2, 0)
    with pytest.raises(lib.NsgError, match="nu and deltat"):
        dev.set_params(nu=-1.0)
    dev.set_params(nu=0.01)
    # zero right-hand side: converged at step 0, x unchanged (SolverControl success on entry)
    dev.set_solution(np.zeros(d.n))
    dev.set_params(p_out=0.0)
    dev.assemble()
    its, res, rc = dev.solve(0, 1e-2, 100, 30, 0)
    assert (its, rc) == (0, 0) and res == 0.0 and not dev.get_delta().any()
    dev.close()


def test_output_files_from_device_solution(pkg, tmp_path):
    """N2: output() (cpp:681-728) after every time step: output-NNNN.xdmf + raw heavy data whose nodal values are the
    device solution at the cell vertices."""
    import importlib
    xout = importlib.import_module("navier-stokes-dealii_b200.output")
    prm = pkg.Parameters(mesh_path=mesh_path("square_h0.1.msh"), nu=0.05, H=1.0, inlet_time_mode="constant", neumann_id=1,
                         inlet_id=0, wall_ids=(2, 3), clear_inlet_before_walls=True, preconditioner="identity", p_out=0.0,
                         increment_bc="consistent", newton_max_iters=3, output_dir=str(tmp_path))
    s = pkg.NavierStokesSolver(2, 1, 0.1, 0.05, prm, verbose=False)
    s.setup()
    s.solve()
    import os
    names = sorted(f for f in os.listdir(tmp_path) if f.endswith(".xdmf"))
    assert names == ["output-0000.xdmf", "output-0001.xdmf", "output-0002.xdmf"]
    t, grids = xout.read_back(str(tmp_path), names[-1])
    assert abs(t - 0.1) < 1e-12
    r = grids["rank0"]
    sol = s.dev.get_solution()
    cd = s.part.cell_dofs.reshape(-1, 15)
    assert r["cells"].shape[0] == s.mesh.n_cells
    assert np.array_equal(r["velocity"][:, 0], sol[cd[:, [0, 3, 6]].reshape(-1)])
    assert np.array_equal(r["pressure"], sol[cd[:, [2, 5, 8]].reshape(-1)])
    assert np.abs(r["velocity"][:, :2]).max() > 0.1
    s.dev.close()
