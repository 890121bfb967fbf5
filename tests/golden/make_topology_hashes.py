"""Pins everything libnst.so produces (mesh numbering by first appearance, red refinement incl. boundary snapping, DoF maps,
partitions, ghost layers, local patterns, halo plans, Dirichlet lists) as SHA-1 digests of the raw arrays for a set of cases.

    python tests/golden/make_topology_hashes.py tests/golden/topology_hashes.json     # regenerate (only after a DELIBERATE change)

tests/test_topology.py::test_topology_outputs_are_pinned recomputes the digests with 1 thread and with all threads and compares:
the numbering must depend neither on the OpenMP thread count nor on how libnst.so computes it (the round-2 rewrite of the serial
"number by first appearance" loops as parallel min/scan passes was checked against the serial library with exactly this file)."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
CASES = [("cylinder_cmy.msh", -1, 0), ("cylinder_cmy.msh", -1, 2), ("cylinder_mesh2d.msh", 5, 3), ("square_h0.05.msh", -1, 1),
         ("square_h0.1.msh", -1, 0)]


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha1(a.tobytes()).hexdigest()[:16] + f":{a.dtype}:{a.shape}"


def compute():
    pkg = importlib.import_module("navier-stokes-dealii_b200")
    golden = os.path.join(ROOT, "tests", "golden")
    h, out = digest, {}
    for name, ent, lv in CASES:
        m = pkg.Mesh.read_msh(os.path.join(golden, name), ent)
        if "mesh2d" in name:
            m.tag_boundary_box(0, 1, 2, 3)
        if lv:
            m = m.refine(lv)
        key = f"{name}:L{lv}"
        out[key + ":mesh"] = [h(m.xy), h(m.cells), h(m.cell_edges), h(m.edge_vertices), h(m.edge_tag), [h(x) for x in m.boundary_faces()]]
        calls = [{11: True}, {11: True, 12: False, 13: False}] if "cmy" in name else [{0: True}, {2: False, 3: False}]
        for parts in (1, 3):
            cp = m.partition_rcb(parts) if parts > 1 else None
            d = pkg.Dofs(m, parts, cp)
            out[key + f":P{parts}:dofs"] = [h(d.cell_dofs), h(d.vertex_node), h(d.edge_node), h(d.vertex_p), int(d.n), int(d.n_u),
                                            h(d.support_points()), h(cp) if cp is not None else None]
            gd, gv = d.dirichlet_values(calls, dict(time_factor=1.0, u_m=1.5, H=1.0))
            out[key + f":P{parts}:dirichlet"] = [h(gd), h(gv)]
            for rank in range(parts):
                for pat in (True, False):
                    p = pkg.Part(d, rank, patterns=pat)
                    out[key + f":P{parts}:r{rank}:pat{int(pat)}"] = [
                        h(p.l2g), h(p.cell_ids), h(p.cell_dofs), h(p.cell_vertices), h(p.xy), h(p.cell_owned), h(p.jac_rowptr), h(p.jac_col),
                        h(p.pm_rowptr), h(p.pm_col), h(p.neighbors), h(p.send_ptr), h(p.send_idx), h(p.recv_ptr), h(p.recv_idx),
                        h(p.bface_cell), h(p.bface_face), h(p.bface_tag)]
    # raw arrays with an inverted cell and a line that is not an edge; refinement with the cylinder boundary snapped to its circle
    xy = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [2, 0.5]], float)
    cells = np.array([[0, 2, 1], [1, 2, 3], [1, 3, 4]], np.int32)
    m = pkg.Mesh.from_arrays(xy, cells, np.array([[0, 1], [3, 4]], np.int32), np.array([7, 9], np.int32)).refine(2)
    out["arrays"] = [h(m.xy), h(m.cells), h(m.cell_edges), h(m.edge_vertices), h(m.edge_tag)]
    m = pkg.Mesh.read_msh(os.path.join(golden, "cylinder_cmy.msh"), -1).refine(3, 13, 0.2, 0.2, 0.05)
    out["snap"] = [h(m.xy), h(m.cells), h(m.cell_edges), h(m.edge_vertices), h(m.edge_tag), [h(x) for x in m.boundary_faces()]]
    d = pkg.Dofs(m)
    out["snap:dofs"] = [h(d.cell_dofs), h(d.support_points())]
    return out


if __name__ == "__main__":
    json.dump(compute(), open(sys.argv[1], "w"), indent=0, sort_keys=True)
