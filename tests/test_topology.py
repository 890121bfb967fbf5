"""Host topology (libnst.so) against the independent structural facts of SURVEY §8d."""
import numpy as np
import pytest

from conftest import MESH_FACTS, mesh_path


@pytest.mark.parametrize("key", list(MESH_FACTS))
def test_counts_match_survey(pkg, key):
    name, ent = key
    f = MESH_FACTS[key]
    m = pkg.Mesh.read_msh(mesh_path(name), ent)
    assert (m.n_vertices, m.n_cells, m.n_edges) == (f["V"], f["T"], f["E"])
    # Euler characteristic of a planar domain with h holes: V - E + T = 1 - h
    holes = 1 if "cylinder" in name else 0
    assert m.n_vertices - m.n_edges + m.n_cells == 1 - holes
    d = pkg.Dofs(m)
    assert (d.n_u, d.n_p) == (f["n_u"], f["n_p"])
    assert d.n_u == 2 * (m.n_vertices + m.n_edges) and d.n_p == m.n_vertices
    if f["nnz"]:
        for kind in range(3):
            rp, col = d.sparsity(kind)
            assert rp[-1] == f["nnz"][kind] == len(col)


def test_overlapping_entities_rejected(pkg):
    # mesh2d.msh as a whole has edges shared by 3 triangles
    with pytest.raises(Exception, match="more than two cells"):
        pkg.Mesh.read_msh(mesh_path("cylinder_mesh2d.msh"), -1)


def test_missing_file_and_bad_args(pkg):
    with pytest.raises(Exception, match="cannot open"):
        pkg.Mesh.read_msh("/nonexistent.msh")


def test_orientation_fixed(pkg):
    m = pkg.Mesh.read_msh(mesh_path("cylinder_mesh2d.msh"), 5)
    assert m.n_inverted == 288  # all clockwise in the file
    xy, c = m.xy, m.cells
    a, b, cc = xy[c[:, 0]], xy[c[:, 1]], xy[c[:, 2]]
    det = (b[:, 0] - a[:, 0]) * (cc[:, 1] - a[:, 1]) - (cc[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    assert (det > 0).all()


def test_boundary_ids_cmy(pkg):
    m = pkg.Mesh.read_msh(mesh_path("cylinder_cmy.msh"))
    _, _, tag = m.boundary_faces()
    counts = {int(t): int((tag == t).sum()) for t in np.unique(tag)}
    assert counts == {10: 40, 11: 40, 12: 140, 13: 32}


def test_dof_numbering_first_visit(pkg):
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    d = pkg.Dofs(m)
    cd = d.cell_dofs
    # replay distribute_dofs + component_wise in pure Python
    node_of, p_of, nn, npp = {}, {}, 0, 0
    cells, ce = m.cells, m.cell_edges
    for c in range(m.n_cells):
        for k in range(3):
            v = ("v", int(cells[c, k]))
            if v not in node_of:
                node_of[v] = nn
                nn += 1
                p_of[v] = npp
                npp += 1
        for k in range(3):
            e = ("e", int(ce[c, k]))
            if e not in node_of:
                node_of[e] = nn
                nn += 1
    n_u = 2 * nn
    for c in range(m.n_cells):
        exp = []
        for k in range(3):
            v = ("v", int(cells[c, k]))
            exp += [2 * node_of[v], 2 * node_of[v] + 1, n_u + p_of[v]]
        for k in range(3):
            e = ("e", int(ce[c, k]))
            exp += [2 * node_of[e], 2 * node_of[e] + 1]
        assert list(cd[c]) == exp


def test_sparsity_is_all_cell_pairs(pkg):
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    d = pkg.Dofs(m)
    cd = d.cell_dofs
    n_u = d.n_u
    sets = [[set() for _ in range(d.n)] for _ in range(3)]
    for c in range(m.n_cells):
        for i in cd[c]:
            for j in cd[c]:
                pp = i >= n_u and j >= n_u
                sets[0][i].add(j)
                if not pp:
                    sets[1][i].add(j)
                else:
                    sets[2][i].add(j)
    for kind in range(3):
        rp, col = d.sparsity(kind)
        for i in range(d.n):
            assert list(col[rp[i]:rp[i + 1]]) == sorted(sets[kind][i])


def test_refinement(pkg):
    m = pkg.Mesh.read_msh(mesh_path("cylinder_cmy.msh"))
    r = m.refine(2)
    assert r.n_cells == 16 * m.n_cells
    assert r.n_vertices - r.n_edges + r.n_cells == 0
    assert r.n_boundary_edges == 4 * m.n_boundary_edges
    _, _, t0 = m.boundary_faces()
    _, _, t2 = r.boundary_faces()
    for t in np.unique(t0):
        assert (t2 == t).sum() == 4 * (t0 == t).sum()
    # area is preserved and children keep the orientation
    def area(mm):
        xy, c = mm.xy, mm.cells
        a, b, cc = xy[c[:, 0]], xy[c[:, 1]], xy[c[:, 2]]
        return 0.5 * ((b[:, 0] - a[:, 0]) * (cc[:, 1] - a[:, 1]) - (cc[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    assert (area(r) > 0).all()
    np.testing.assert_allclose(area(r).sum(), area(m).sum(), rtol=1e-13)


def test_dirichlet_list(pkg):
    m = pkg.Mesh.read_msh(mesh_path("cylinder_cmy.msh"))
    d = pkg.Dofs(m)
    inlet = dict(u_m=1.5, H=0.41, time_factor=1.0)
    gd, gv = d.dirichlet_values([{11: True}, {11: True, 12: False, 13: False}], inlet)
    assert len(gd) == 850 and (np.diff(gd) > 0).all() and gd.max() < d.n_u
    xy = d.support_points()
    on_inlet = np.isclose(xy[gd, 0], 5.0)
    ux = (gd % 2 == 0)
    y = xy[gd, 1]
    exp = np.where(on_inlet & ux, 4 * 1.5 * y * (0.41 - y) / 0.41 ** 2, 0.0)
    # corners shared by the inlet and a wall take whichever face is visited last; everything else is exact
    interior = ~(on_inlet & (np.isclose(np.abs(y), 2.0)))
    np.testing.assert_allclose(gv[interior], exp[interior], rtol=0, atol=1e-14)
    # frozen time: all values are zero
    _, gv0 = d.dirichlet_values([{11: True}, {11: True, 12: False, 13: False}], dict(u_m=1.5, H=0.41, time_factor=0.0))
    assert (gv0 == 0).all()


@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_partition_and_local_problems(pkg, n_parts):
    m = pkg.Mesh.read_msh(mesh_path("cylinder_cmy.msh"))
    cp = m.partition_rcb(n_parts)
    sizes = np.bincount(cp, minlength=n_parts)
    assert sizes.min() > 0 and sizes.max() - sizes.min() <= n_parts
    d1 = pkg.Dofs(m)
    d = pkg.Dofs(m, n_parts, cp)
    assert (d.n_u, d.n_p) == (d1.n_u, d1.n_p)
    assert np.array_equal(np.unique(d.cell_dofs), np.arange(d.n))
    rp, col = d.sparsity(0)
    parts = [pkg.Part(d, r) for r in range(n_parts)]
    owned = np.concatenate([p.l2g[:p.n_own] for p in parts])
    assert sorted(owned.tolist()) == list(range(d.n))          # every DoF owned exactly once
    assert sum(p.n_owned_cells for p in parts) == m.n_cells
    for p in parts:
        # block-wise, part-major numbering: owned ranges are contiguous per block
        assert (np.diff(p.l2g[:p.n_own_u]) == 1).all() and (np.diff(p.l2g[p.n_own_u:p.n_own]) == 1).all()
        # the local pattern is the global pattern of the owned rows, in local column ids
        for i in np.random.default_rng(p.rank).integers(0, p.n_own, 50):
            g = p.l2g[i]
            loc = p.l2g[p.jac_col[p.jac_rowptr[i]:p.jac_rowptr[i + 1]]]
            assert sorted(loc.tolist()) == col[rp[g]:rp[g + 1]].tolist()
            assert (np.diff(p.jac_col[p.jac_rowptr[i]:p.jac_rowptr[i + 1]]) > 0).all()
    # halo plans are mutually consistent: what r sends to k is what k expects from r, same order
    for p in parts:
        for a, k in enumerate(p.neighbors):
            q = parts[k]
            b = list(q.neighbors).index(p.rank)
            sent = p.l2g[p.send_idx[p.send_ptr[a]:p.send_ptr[a + 1]]]
            recv = q.l2g[q.recv_idx[q.recv_ptr[b]:q.recv_ptr[b + 1]]]
            assert sent.tolist() == recv.tolist() and len(sent) > 0
        assert p.n_recv == p.n_ghost_u + p.n_ghost_p


def test_output_writer_xdmf_roundtrip(pkg, tmp_path):
    """N2: the XDMF + raw-binary writer (cpp:681-728 without HDF5): one 3-node patch per owned cell with velocity,
    pressure and partitioning, two ranks described by one .xdmf; read back through the XML."""
    import importlib
    xout = importlib.import_module("navier-stokes-dealii_b200.output")
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    cp = m.partition_rcb(2)
    d = pkg.Dofs(m, 2, cp)
    xy = d.support_points()
    g = np.zeros(d.n)
    g[0:d.n_u:2] = 1.0 + xy[0:d.n_u:2, 0]           # u_x = 1 + x
    g[1:d.n_u:2] = 2.0 * xy[1:d.n_u:2, 1]           # u_y = 2 y
    g[d.n_u:] = xy[d.n_u:, 0] - xy[d.n_u:, 1]       # p = x - y
    layouts, n_cells = {}, 0
    for rank in (0, 1):
        part = pkg.Part(d, rank)
        patches = xout.cell_patches(part, g[part.l2g], rank)
        layouts[rank] = xout.write_rank_file(str(tmp_path), "output-0003", rank, patches)
        n_cells += int(part.cell_owned.sum())
    assert n_cells == m.n_cells                      # every cell is written by exactly one rank
    xout.write_xdmf(str(tmp_path), "output-0003", 0.15, layouts)
    t, grids = xout.read_back(str(tmp_path), "output-0003.xdmf")
    assert t == 0.15 and sorted(grids) == ["rank0", "rank1"]
    for rank in (0, 1):
        r = grids[f"rank{rank}"]
        T = r["cells"].shape[0]
        assert r["points"].shape == (3 * T, 2) and np.array_equal(r["cells"].reshape(-1), np.arange(3 * T))
        x, y = r["points"][:, 0], r["points"][:, 1]
        assert np.allclose(r["velocity"][:, 0], 1.0 + x, atol=1e-14) and np.allclose(r["velocity"][:, 1], 2.0 * y, atol=1e-14)
        assert not r["velocity"][:, 2].any() and np.allclose(r["pressure"], x - y, atol=1e-14)
        assert (r["partitioning"] == rank).all()


def test_part_arrays_are_views_kept_alive_by_copies(pkg):
    """The big Part arrays are read-only zero-copy views of library-owned buffers; a shallow copy of a Part (used by the
    tests to swap in modified patterns) must keep those buffers alive after the original is gone."""
    import copy
    import gc
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    d = pkg.Dofs(m)
    part = pkg.Part(d, 0)
    ref_col, ref_dofs = part.jac_col.copy(), part.cell_dofs.copy()
    assert not part.jac_col.flags.writeable and not part.cell_dofs.flags.writeable
    with pytest.raises(ValueError):
        part.jac_col[0] = 7
    clone = copy.copy(part)
    del part
    gc.collect()
    junk = [np.ones(1 << 16) for _ in range(64)]      # churn the allocator: freed memory would be overwritten
    assert np.array_equal(clone.jac_col, ref_col) and np.array_equal(clone.cell_dofs, ref_dofs)
    del junk


def test_part_views_outlive_the_part(pkg):
    """The big Part arrays are zero-copy views of library-owned buffers: a slice that outlives its Part must keep the buffers
    alive (it holds the handle), and a Part built without the sparsity patterns (they are built on the device) has none."""
    import gc
    m = pkg.Mesh.read_msh(mesh_path("square_h0.1.msh"))
    d = pkg.Dofs(m)
    part = pkg.Part(d, 0)
    keep = part.jac_col[5:50]
    want = np.array(keep)
    cd = np.array(part.cell_dofs)
    lean = pkg.Part(d, 0, patterns=False)
    assert not lean.has_patterns and lean.nnz_jac == 0 and lean.nnz_pm == 0 and not lean.jac_rowptr.any()
    assert np.array_equal(lean.cell_dofs, cd) and lean.n_own == part.n_own and np.array_equal(lean.l2g, part.l2g)
    del part
    gc.collect()
    junk = [np.zeros(1 << 16) for _ in range(8)]      # churn the allocator
    assert np.array_equal(keep, want) and keep._owner is not None
    del junk


@pytest.mark.parametrize("threads", ["1", "all"])
def test_topology_outputs_are_pinned(threads):
    """Everything libnst.so produces (numbering by first appearance of lines, vertices and DoFs, red refinement incl. snapping, partitions,
    ghost layers, local patterns, halo plans, Dirichlet lists) equals the committed digests - which were generated with the serial
    round-1 library - for one OpenMP thread and for all of them: the parallel min/scan numbering is deterministic and unchanged."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = os.path.join(root, "tests", "golden", "make_topology_hashes.py")
    env = dict(os.environ)
    if threads == "1":
        env["OMP_NUM_THREADS"] = "1"
    else:
        env.pop("OMP_NUM_THREADS", None)
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"topology_hashes_{os.getpid()}_{threads}.json")
    try:
        subprocess.run([sys.executable, script, out], check=True, env=env, timeout=600)
        got, want = json.load(open(out)), json.load(open(os.path.join(root, "tests", "golden", "topology_hashes.json")))
    finally:
        if os.path.exists(out):
            os.remove(out)
    assert got.keys() == want.keys()
    assert [k for k in want if got[k] != want[k]] == []
