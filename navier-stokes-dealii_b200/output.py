"""N2: NavierStokesSolver::output (src/NavierStokesSolver.cpp:681-728).

The reference hands the ghosted `solution` to DataOut with the default one subdivision per cell and
`filter_duplicate_vertices = false`, i.e. every cell becomes a patch of its three vertices carrying the
nodal values of velocity (a vector), pressure and the cell's subdomain id ("partitioning"), and writes
them as XDMF + HDF5.  HDF5 is not in this image, so the heavy data go into one raw little-endian binary
file per rank (XDMF `Format="Binary"` with `Seek` offsets) described by the same kind of XDMF file:
`output-NNNN.xdmf` (written by rank 0, one sub-grid per rank) + `output-NNNN.rankR.bin`."""
import os

import numpy as np


def cell_patches(part, solution_ghosted, rank):
    """Arrays DataOut::build_patches would produce for this rank's OWNED cells.
    Returns dict(points (3T,2), cells (T,3) int32, velocity (3T,3), pressure (3T,), partitioning (3T,))."""
    own = np.asarray(part.cell_owned).astype(bool)
    cd = np.asarray(part.cell_dofs).reshape(-1, 15)[own]
    cv = np.asarray(part.cell_vertices).reshape(-1, 3)[own]
    T = len(cd)
    sol = np.asarray(solution_ghosted, np.float64)
    pts = np.asarray(part.xy).reshape(-1, 2)[cv.reshape(-1)]
    vel = np.zeros((3 * T, 3))                       # XDMF vectors are 3-component; u_z = 0
    vel[:, 0] = sol[cd[:, [0, 3, 6]].reshape(-1)]    # FESystem order: vertex v -> (u_x, u_y, p) at 3v, 3v+1, 3v+2
    vel[:, 1] = sol[cd[:, [1, 4, 7]].reshape(-1)]
    pres = sol[cd[:, [2, 5, 8]].reshape(-1)]
    return {"points": pts, "cells": np.arange(3 * T, dtype=np.int32).reshape(T, 3), "velocity": vel, "pressure": pres,
            "partitioning": np.full(3 * T, float(rank))}


def write_rank_file(directory, name, rank, patches):
    """Heavy data of one rank; returns the layout [(key, offset, shape, number type, precision)]."""
    path = os.path.join(directory, f"{name}.rank{rank}.bin")
    layout, off = [], 0
    with open(path, "wb") as f:
        for key in ("points", "cells", "velocity", "pressure", "partitioning"):
            a = np.ascontiguousarray(patches[key])
            a = a.astype("<i4") if a.dtype.kind == "i" else a.astype("<f8")
            f.write(a.tobytes())
            layout.append((key, off, a.shape, "Int" if a.dtype.kind == "i" else "Float", a.dtype.itemsize))
            off += a.nbytes
    return layout


def _item(fname, entry):
    key, off, shape, ntype, prec = entry
    dims = " ".join(str(int(x)) for x in shape)
    return (f'<DataItem Dimensions="{dims}" NumberType="{ntype}" Precision="{prec}" Format="Binary" Endian="Little" '
            f'Seek="{off}">{fname}</DataItem>')


def write_xdmf(directory, name, time, layouts):
    """layouts: {rank: layout from write_rank_file}.  One uniform grid per rank inside a spatial collection."""
    lines = ['<?xml version="1.0" ?>', '<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>', '<Xdmf Version="3.0">', " <Domain>",
             '  <Grid Name="CellTime" GridType="Collection" CollectionType="Temporal">',
             '   <Grid Name="mesh" GridType="Collection" CollectionType="Spatial">', f'    <Time Value="{time!r}"/>']
    for rank in sorted(layouts):
        lay = {e[0]: e for e in layouts[rank]}
        fname = f"{name}.rank{rank}.bin"
        n_cells = lay["cells"][2][0]
        lines += [f'    <Grid Name="rank{rank}" GridType="Uniform">',
                  f'     <Topology TopologyType="Triangle" NumberOfElements="{n_cells}">{_item(fname, lay["cells"])}</Topology>',
                  f'     <Geometry GeometryType="XY">{_item(fname, lay["points"])}</Geometry>',
                  f'     <Attribute Name="velocity" AttributeType="Vector" Center="Node">{_item(fname, lay["velocity"])}</Attribute>',
                  f'     <Attribute Name="pressure" AttributeType="Scalar" Center="Node">{_item(fname, lay["pressure"])}</Attribute>',
                  f'     <Attribute Name="partitioning" AttributeType="Scalar" Center="Node">{_item(fname, lay["partitioning"])}'
                  "</Attribute>", "    </Grid>"]
    lines += ["   </Grid>", "  </Grid>", " </Domain>", "</Xdmf>"]
    path = os.path.join(directory, f"{name}.xdmf")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def read_back(directory, xdmf_name):
    """Parse an .xdmf written above and load its binary items (used by the tests; also a usage example)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(os.path.join(directory, xdmf_name)).getroot()
    out = {}
    for grid in root.iter("Grid"):
        if grid.get("GridType") != "Uniform":
            continue
        rec = {}
        for node, key in ((grid.find("Topology"), "cells"), (grid.find("Geometry"), "points")):
            rec[key] = _load(directory, node.find("DataItem"))
        for att in grid.findall("Attribute"):
            rec[att.get("Name")] = _load(directory, att.find("DataItem"))
        out[grid.get("Name")] = rec
    t = next(root.iter("Time"))
    return float(t.get("Value")), out


def _load(directory, item):
    shape = tuple(int(x) for x in item.get("Dimensions").split())
    dt = ("<i" if item.get("NumberType") == "Int" else "<f") + item.get("Precision")
    n = int(np.prod(shape))
    return np.fromfile(os.path.join(directory, item.text.strip()), dtype=dt, count=n, offset=int(item.get("Seek"))).reshape(shape)
