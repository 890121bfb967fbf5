"""Host topology objects over libnst.so (include/nst.h): what NavierStokesSolver::setup()
(reference src/NavierStokesSolver.cpp:4-176) produces for the hot path."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import InletParams, PartInfo, as_array, nst, nst_check


class Mesh:
    """Triangulation read by GridIn::read_msh (cpp:12-16) or built from arrays."""

    def __init__(self, handle):
        self._h = handle
        L = nst()
        self.n_vertices = L.nst_mesh_n_vertices(handle)
        self.n_cells = L.nst_mesh_n_cells(handle)
        self.n_edges = L.nst_mesh_n_edges(handle)
        self.n_boundary_edges = L.nst_mesh_n_boundary_edges(handle)
        self.n_inverted = L.nst_mesh_n_inverted(handle)

    @classmethod
    def read_msh(cls, path, surface_entity=-1):
        h = C.c_void_p()
        nst_check(nst().nst_mesh_read_msh(str(path).encode(), int(surface_entity), C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, xy, cells, line_v=None, line_tag=None):
        xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1)
        cells = np.ascontiguousarray(cells, dtype=np.int32).reshape(-1)
        nl = 0
        if line_v is not None:
            line_v = np.ascontiguousarray(line_v, dtype=np.int32).reshape(-1)
            line_tag = np.ascontiguousarray(line_tag, dtype=np.int32).reshape(-1)
            nl = len(line_tag)
        h = C.c_void_p()
        nst_check(nst().nst_mesh_create(len(xy) // 2, xy, len(cells) // 3, cells, nl, line_v, line_tag, C.byref(h)))
        return cls(h)

    def refine(self, levels, snap_id=-1, cx=0.0, cy=0.0, r=0.0):
        h = C.c_void_p()
        nst_check(nst().nst_mesh_refine(self._h, int(levels), int(snap_id), cx, cy, r, C.byref(h)))
        return Mesh(h)

    def tag_boundary_box(self, id_left, id_right, id_wall, id_other):
        nst_check(nst().nst_mesh_tag_boundary_box(self._h, id_left, id_right, id_wall, id_other))

    @property
    def xy(self):
        return as_array(nst().nst_mesh_xy(self._h), 2 * self.n_vertices, np.float64).reshape(-1, 2)

    @property
    def cells(self):
        return as_array(nst().nst_mesh_cells(self._h), 3 * self.n_cells, np.int32).reshape(-1, 3)

    @property
    def cell_edges(self):
        return as_array(nst().nst_mesh_cell_edges(self._h), 3 * self.n_cells, np.int32).reshape(-1, 3)

    @property
    def edge_vertices(self):
        return as_array(nst().nst_mesh_edge_vertices(self._h), 2 * self.n_edges, np.int32).reshape(-1, 2)

    @property
    def edge_tag(self):
        return as_array(nst().nst_mesh_edge_tag(self._h), self.n_edges, np.int32)

    def boundary_faces(self):
        n = self.n_boundary_edges
        c, f, t = (np.zeros(n, np.int32) for _ in range(3))
        nst_check(nst().nst_mesh_boundary_faces(self._h, c, f, t))
        return c, f, t

    def partition_rcb(self, n_parts):
        part = np.zeros(self.n_cells, np.int32)
        nst_check(nst().nst_partition_rcb(self._h, int(n_parts), part))
        return part

    def __del__(self):
        try:
            if self._h:
                nst().nst_mesh_free(self._h)
        except Exception:
            pass


class Dofs:
    """DoFHandler::distribute_dofs + component_wise renumbering (cpp:64-91)."""

    def __init__(self, mesh, n_parts=1, cell_part=None):
        self.mesh = mesh
        self.n_parts = int(n_parts)
        self.cell_part = None if cell_part is None else np.ascontiguousarray(cell_part, dtype=np.int32)
        h = C.c_void_p()
        nst_check(nst().nst_dofs_distribute(mesh._h, self.n_parts, self.cell_part, C.byref(h)))
        self._h = h
        L = nst()
        self.n_u = L.nst_dofs_n_u(h)
        self.n_p = L.nst_dofs_n_p(h)
        self.n = self.n_u + self.n_p
        self.part_n_u = as_array(L.nst_dofs_part_n_u(h), self.n_parts, np.int64)
        self.part_n_p = as_array(L.nst_dofs_part_n_p(h), self.n_parts, np.int64)

    @property
    def cell_dofs(self):
        return as_array(nst().nst_dofs_cell_dofs(self._h), 15 * self.mesh.n_cells, np.int32).reshape(-1, 15)

    @property
    def vertex_node(self):
        return as_array(nst().nst_dofs_vertex_node(self._h), self.mesh.n_vertices, np.int32)

    @property
    def edge_node(self):
        return as_array(nst().nst_dofs_edge_node(self._h), self.mesh.n_edges, np.int32)

    @property
    def vertex_p(self):
        return as_array(nst().nst_dofs_vertex_p(self._h), self.mesh.n_vertices, np.int32)

    def sparsity(self, kind):
        """kind 0: Jacobian (cpp:107-110), 1: Stokes (cpp:124-140), 2: pressure mass (cpp:143-158)."""
        nnz = C.c_int64()
        nst_check(nst().nst_sparsity(self.mesh._h, self._h, kind, C.byref(nnz), None, None))
        rowptr = np.zeros(self.n + 1, np.int64)
        col = np.zeros(max(nnz.value, 1), np.int32)
        nst_check(nst().nst_sparsity(self.mesh._h, self._h, kind, C.byref(nnz), rowptr, col))
        return rowptr, col[: nnz.value]

    def support_points(self):
        xy = np.empty(2 * self.n, np.float64)   # every entry is written
        nst_check(nst().nst_dofs_support_points(self.mesh._h, self._h, xy))
        return xy.reshape(-1, 2)

    def dirichlet_values(self, calls, inlet):
        """calls: list of dicts {boundary_id: is_inlet(bool)} = the successive
        interpolate_boundary_values calls (cpp:357-373). Returns (dofs, values), global ids."""
        ptr, ids, isin = [0], [], []
        for call in calls:
            for k, v in call.items():
                ids.append(int(k))
                isin.append(1 if v else 0)
            ptr.append(len(ids))
        ptr = np.asarray(ptr, np.int32)
        ids = np.asarray(ids, np.int32)
        isin = np.asarray(isin, np.int32)
        ip = InletParams(inlet["u_m"], inlet["H"], inlet.get("y0", 0.0), inlet["time_factor"])
        n = C.c_int64(0)
        nst_check(nst().nst_dirichlet_values(self.mesh._h, self._h, len(calls), ptr, ids, isin, C.byref(ip), C.byref(n),
                                             None, None))
        dofs = np.zeros(max(n.value, 1), np.int32)
        vals = np.zeros(max(n.value, 1), np.float64)
        nst_check(nst().nst_dirichlet_values(self.mesh._h, self._h, len(calls), ptr, ids, isin, C.byref(ip), C.byref(n),
                                             dofs, vals))
        return dofs[: n.value], vals[: n.value]

    def __del__(self):
        try:
            if self._h:
                nst().nst_dofs_free(self._h)
        except Exception:
            pass


class _PartHandle:
    """Owns one nst_part; freed when the last Part (or shallow copy of it) that views its buffers goes away."""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        try:
            if self._h:
                nst().nst_part_free(self._h)
                self._h = None
        except Exception:
            pass


class _OwnedView(np.ndarray):
    """ndarray over a library-owned buffer that keeps the owner of the buffer alive: every view / slice derived from it
    inherits the reference (`__array_finalize__`), so an array that outlives its Part does not read freed memory."""

    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


def _view(ptr, n, dtype, owner=None):
    """Read-only zero-copy view of a library-owned buffer (empty array for n == 0); `owner` is kept alive by the view."""
    if n == 0:
        return np.zeros(0, dtype=dtype)
    a = np.ctypeslib.as_array(ptr, shape=(int(n),))
    assert a.dtype == np.dtype(dtype), (a.dtype, dtype)
    a.flags.writeable = False
    v = a.view(_OwnedView)
    v._owner = owner
    return v


class Part:
    """One rank's local problem: owned rows + one ghost layer (cpp:19-21, 75-91)."""

    def __init__(self, dofs, rank=0, patterns=True):
        """patterns=False: leave the two sparsity patterns to the device (DeviceProblem builds them from cell_dofs)."""
        self.dofs = dofs
        self.rank = int(rank)
        self.has_patterns = bool(patterns)
        h = C.c_void_p()
        nst_check(nst().nst_part_build_ex(dofs.mesh._h, dofs._h, dofs.n_parts, dofs.cell_part, self.rank, 0 if patterns else 1,
                                          C.byref(h)))
        self._h = h
        info = PartInfo()
        nst_check(nst().nst_part_get_info(h, C.byref(info)))
        for name, _ in PartInfo._fields_:
            setattr(self, name, int(getattr(info, name)))
        self.n_own = self.n_own_u + self.n_own_p
        self.n_loc = self.n_own + self.n_ghost_u + self.n_ghost_p
        L = nst()
        # the big arrays are VIEWS of the library-owned buffers (no second copy of hundreds of millions of pattern
        # entries); the holder keeps the nst_part alive for as long as any (shallow copy of this) Part refers to it
        self._keep = _PartHandle(h)

        def view(ptr, n, dtype):
            return _view(ptr, n, dtype, self._keep)
        self.l2g = view(L.nst_part_l2g(h), self.n_loc, np.int64)
        self.cell_ids = view(L.nst_part_cell_ids(h), self.n_cells, np.int32)
        self.cell_dofs = view(L.nst_part_cell_dofs(h), 15 * self.n_cells, np.int32)
        self.cell_vertices = view(L.nst_part_cell_vertices(h), 3 * self.n_cells, np.int32)
        self.xy = view(L.nst_part_xy(h), 2 * self.n_vertices, np.float64)
        self.cell_owned = view(L.nst_part_cell_owned(h), self.n_cells, np.uint8)
        if self.has_patterns:
            self.jac_rowptr = view(L.nst_part_jac_rowptr(h), self.n_own + 1, np.int64)
            self.pm_rowptr = view(L.nst_part_pm_rowptr(h), self.n_own + 1, np.int64)
        else:   # the library keeps no row pointers in this mode; all-zero ones (lazily allocated pages) keep the attribute
            self.jac_rowptr, self.pm_rowptr = np.zeros(self.n_own + 1, np.int64), np.zeros(self.n_own + 1, np.int64)
        self.jac_col = view(L.nst_part_jac_col(h), self.nnz_jac, np.int32)
        self.pm_col = view(L.nst_part_pm_col(h), self.nnz_pm, np.int32)
        self.neighbors = as_array(L.nst_part_neighbors(h), self.n_neighbors, np.int32)
        self.send_ptr = as_array(L.nst_part_send_ptr(h), self.n_neighbors + 1, np.int64)
        self.send_idx = as_array(L.nst_part_send_idx(h), self.n_send, np.int32)
        self.recv_ptr = as_array(L.nst_part_recv_ptr(h), self.n_neighbors + 1, np.int64)
        self.recv_idx = as_array(L.nst_part_recv_idx(h), self.n_recv, np.int32)
        nb = L.nst_part_n_boundary_faces(h)
        self.bface_cell = as_array(L.nst_part_bface_cell(h), nb, np.int32)
        self.bface_face = as_array(L.nst_part_bface_face(h), nb, np.int32)
        self.bface_tag = as_array(L.nst_part_bface_tag(h), nb, np.int32)
        self._h = None
        # global -> local map of the owned rows (for the Dirichlet list)
        self.own_global = self.l2g[: self.n_own]

    def localize_dirichlet(self, gdofs, gvals):
        """Keep the entries this rank owns and translate them to local row ids."""
        u0, p0 = (int(self.own_global[0]) if self.n_own_u else 0), None
        g = np.asarray(gdofs, np.int64)
        n_u = self.dofs.n_u
        lo_u = int(self.dofs.part_n_u[: self.rank].sum())
        lo_p = int(self.dofs.part_n_p[: self.rank].sum())
        is_u = (g < n_u) & (g >= lo_u) & (g < lo_u + self.n_own_u)
        is_p = (g >= n_u) & (g - n_u >= lo_p) & (g - n_u < lo_p + self.n_own_p)
        loc = np.where(is_u, g - lo_u, self.n_own_u + (g - n_u - lo_p))
        keep = is_u | is_p
        del u0, p0
        return loc[keep].astype(np.int32), np.asarray(gvals, np.float64)[keep].copy()
