"""ctypes wrapper of oracle/libns_oracle.so — the CPU ORACLE (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under navier-stokes-dealii_b200/ does.  See ns_oracle.cpp's header:
PARITY UNPINNED (the reference has no golden vectors and deal.II/Trilinos are absent here).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(_HERE, "libns_oracle.so"))
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i64, i64, i64p, i32p, i64p, i32p, i64, i64, f64p, i32p, i32p, i64, i32p, i32p, i32p]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_params.argtypes = [vp, dbl, dbl, dbl, dbl, dbl, dbl, i32, i32, i32, i32]
        L.orc_set_block_jacobi.argtypes = [vp, C.c_int, i64p, i64p]
        L.orc_assemble.argtypes = [vp]
        L.orc_apply_dirichlet.argtypes = [vp, i64, i32p, f64p, i32]
        L.orc_residual_norm.restype = dbl
        L.orc_residual_norm.argtypes = [vp]
        L.orc_spmv.argtypes = [vp, f64p, f64p]
        L.orc_solve.restype = C.c_int
        L.orc_solve.argtypes = [vp, C.c_int, dbl, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(dbl)]
        L.orc_last_inner_iterations.restype = i64
        L.orc_last_inner_iterations.argtypes = [vp]
        L.orc_gmres_history.restype = i64
        L.orc_gmres_history.argtypes = [vp, f64p, i64]
        for n in ("orc_update_solution", "orc_push_time_level"):
            getattr(L, n).argtypes = [vp]
        for n in ("orc_set_solution", "orc_set_solution_old", "orc_set_delta", "orc_get_solution", "orc_get_delta",
                  "orc_get_residual", "orc_get_matrix_values", "orc_get_pm_values"):
            getattr(L, n).argtypes = [vp, f64p]
        L.orc_ilu_apply.argtypes = [vp, C.c_int, f64p, f64p]
        L.orc_boundary_force.argtypes = [vp, i32, f64p]
        L.orc_quadrature.argtypes = [f64p, f64p, f64p]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _nz(a, dtype=np.int32):
    a = np.ascontiguousarray(a, dtype)
    return a if a.size else np.zeros(1, dtype)


class Oracle:
    """Mirrors DeviceProblem so that parity tests read symmetrically (global numbering, P = 1)."""

    def __init__(self, part):
        L = lib()
        self._L = L
        assert part.n_ghost_u == 0 and part.n_ghost_p == 0, "the oracle works on the undistributed problem"
        self.n = part.n_own
        self.nnz, self.pm_nnz = part.nnz_jac, part.nnz_pm
        self._h = L.orc_create(part.n_own_u, part.n_own_p, part.jac_rowptr, _nz(part.jac_col), part.pm_rowptr,
                               _nz(part.pm_col), part.n_cells, part.n_vertices, part.xy, part.cell_vertices,
                               part.cell_dofs, len(part.bface_cell), _nz(part.bface_cell), _nz(part.bface_face),
                               _nz(part.bface_tag))

    def set_params(self, nu=0.001, rho=1.0, p_out=10.0, deltat=0.05, forcing=(0.0, 0.0), neumann_id=10, use_mass=1,
                   stokes=0, dirichlet_diag=0):
        self._L.orc_set_params(self._h, nu, rho, p_out, deltat, forcing[0], forcing[1], neumann_id, use_mass, stokes,
                               dirichlet_diag)

    def set_block_jacobi(self, u_off, p_off):
        u_off = np.ascontiguousarray(u_off, np.int64)
        p_off = np.ascontiguousarray(p_off, np.int64)
        self._L.orc_set_block_jacobi(self._h, len(u_off) - 1, u_off, p_off)

    def assemble(self):
        self._L.orc_assemble(self._h)

    def apply_dirichlet(self, dofs, values, into_solution=False):
        dofs = np.ascontiguousarray(dofs, np.int32)
        values = np.ascontiguousarray(values, np.float64)
        self._L.orc_apply_dirichlet(self._h, len(dofs), _nz(dofs), _nz(values, np.float64), 1 if into_solution else 0)

    def residual_norm(self):
        return self._L.orc_residual_norm(self._h)

    def spmv(self, x):
        y = np.zeros(self.n)
        self._L.orc_spmv(self._h, np.ascontiguousarray(x, np.float64), y)
        return y

    def solve(self, precond=0, rel_tol=1e-2, max_it=100000, n_tmp=30, target=0):
        its, res = C.c_int(), C.c_double()
        rc = self._L.orc_solve(self._h, precond, rel_tol, max_it, n_tmp, target, C.byref(its), C.byref(res))
        return its.value, res.value, rc

    def last_inner_iterations(self):
        return int(self._L.orc_last_inner_iterations(self._h))

    def gmres_history(self):
        n = self._L.orc_gmres_history(self._h, np.zeros(1), 0)
        out = np.zeros(max(n, 1))
        self._L.orc_gmres_history(self._h, out, n)
        return out[:n]

    def update_solution(self):
        self._L.orc_update_solution(self._h)

    def push_time_level(self):
        self._L.orc_push_time_level(self._h)

    def _get(self, fn, n):
        out = np.zeros(max(n, 1))
        fn(self._h, out)
        return out[:n]

    def set_solution(self, v):
        self._L.orc_set_solution(self._h, np.ascontiguousarray(v, np.float64))

    def set_solution_old(self, v):
        self._L.orc_set_solution_old(self._h, np.ascontiguousarray(v, np.float64))

    def set_delta(self, v):
        self._L.orc_set_delta(self._h, np.ascontiguousarray(v, np.float64))

    def get_solution(self):
        return self._get(self._L.orc_get_solution, self.n)

    def get_solution_ghosted(self):  # one rank: no ghost layer
        return self.get_solution()

    def get_delta(self):
        return self._get(self._L.orc_get_delta, self.n)

    def get_residual(self):
        return self._get(self._L.orc_get_residual, self.n)

    def get_matrix_values(self):
        return self._get(self._L.orc_get_matrix_values, self.nnz)

    def get_pm_values(self):
        return self._get(self._L.orc_get_pm_values, self.pm_nnz)

    def boundary_force(self, boundary_id):
        out = np.zeros(2)
        self._L.orc_boundary_force(self._h, int(boundary_id), out)
        return out

    def ilu_apply(self, which, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(len(x))
        self._L.orc_ilu_apply(self._h, which, x, y)
        return y

    def __del__(self):
        try:
            if self._h:
                self._L.orc_destroy(self._h)
                self._h = None
        except Exception:
            pass


def quadrature():
    x, y, w = np.zeros(7), np.zeros(7), np.zeros(7)
    lib().orc_quadrature(x, y, w)
    return x, y, w
