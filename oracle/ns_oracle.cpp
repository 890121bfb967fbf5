// ns_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY: nothing under navier-stokes-dealii_b200/
// may import, link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker and the reported CPU baseline.
//
// PARITY UNPINNED: the reference (giuseppeegentile/Navier-Stokes-dealii) has no tests or golden
// vectors, and its arithmetic lives in deal.II (>= 9.3.1, common/cmake-common.cmake:28) and
// Trilinos/Ifpack, neither of which exists in this image.  This file restates, in plain C++,
// the reference's own loops (src/NavierStokesSolver.cpp) plus the published algorithms of the
// deal.II / Ifpack routines they call (SolverGMRES, SolverCG, MatrixTools::apply_boundary_values
// for Trilinos block matrices, Ifpack_ILU level 0).  It is pinned only by the substitutes in
// tests/: exact polynomial integrals, structural counts and the u=0,p=10 fixed point.
//
// Reference sites (relative to /root/reference):
//   assemble_system           src/NavierStokesSolver.cpp:178-378
//   assemble_stokes_system    src/NavierStokesSolver.cpp:380-531
//   solve_system              src/NavierStokesSolver.cpp:561-588
//   solve_stokes_system       src/NavierStokesSolver.cpp:533-559
//   solve_newton              src/NavierStokesSolver.cpp:590-627
//   PreconditionIdentity / BlockDiagonal / BlockTriangular  src/NavierStokesSolver.hpp:504-639
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

// --- FE_SimplexP(2)^2 x FE_SimplexP(1), local order: vertex v -> 3v+{0,1} (u), 3v+2 (p);
//     line l -> 9+2l+{0,1}.  QGaussSimplex<2>(3): 7 points, degree 5.  (SURVEY §9-2,3)
struct Quad7 {
  double x[7], y[7], w[7];
  Quad7() {
    const double s = std::sqrt(15.0);
    const double a = (6.0 - s) / 21.0, b = (6.0 + s) / 21.0;
    const double wa = (155.0 - s) / 2400.0, wb = (155.0 + s) / 2400.0;
    const double px[7] = {1.0 / 3.0, 1 - 2 * a, a, a, 1 - 2 * b, b, b};
    const double py[7] = {1.0 / 3.0, a, 1 - 2 * a, a, b, 1 - 2 * b, b};
    const double pw[7] = {9.0 / 80.0, wa, wa, wa, wb, wb, wb};
    for (int q = 0; q < 7; ++q) x[q] = px[q], y[q] = py[q], w[q] = pw[q];
  }
};
const Quad7 Q7;

// scalar P2 shape functions and reference gradients at (x,y)
void p2_eval(double x, double y, double psi[6], double dpsi[6][2]) {
  const double l0 = 1 - x - y, l1 = x, l2 = y;
  psi[0] = l0 * (2 * l0 - 1);
  psi[1] = l1 * (2 * l1 - 1);
  psi[2] = l2 * (2 * l2 - 1);
  psi[3] = 4 * l0 * l1;
  psi[4] = 4 * l1 * l2;
  psi[5] = 4 * l2 * l0;
  const double d0[2] = {-1, -1}, d1[2] = {1, 0}, d2[2] = {0, 1};
  for (int c = 0; c < 2; ++c) {
    dpsi[0][c] = (4 * l0 - 1) * d0[c];
    dpsi[1][c] = (4 * l1 - 1) * d1[c];
    dpsi[2][c] = (4 * l2 - 1) * d2[c];
    dpsi[3][c] = 4 * (l0 * d1[c] + l1 * d0[c]);
    dpsi[4][c] = 4 * (l1 * d2[c] + l2 * d1[c]);
    dpsi[5][c] = 4 * (l2 * d0[c] + l0 * d2[c]);
  }
}

// local dof i -> (is_pressure, component, scalar index)
inline void local_dof(int i, int &is_p, int &comp, int &k) {
  if (i < 9) {
    const int v = i / 3, r = i % 3;
    if (r == 2) {
      is_p = 1, comp = 2, k = v;
    } else {
      is_p = 0, comp = r, k = v;
    }
  } else {
    is_p = 0, comp = (i - 9) % 2, k = 3 + (i - 9) / 2;
  }
}

struct Params {
  double nu, rho, p_out, deltat, f[2];
  int32_t neumann_id;
  int32_t use_mass;  // 1: implicit-Euler mass terms as in cpp:249-251,288; 0: steady
  int32_t stokes;    // 1: assemble_stokes_system (cpp:380-531): viscous + B/Bt only, rhs = forcing + Neumann
  int32_t dirichlet_diag;  // 0: Trilinos rule (diagonal always replaced by the block's first non-zero diagonal); 1: keep a non-zero diagonal
};

struct Ctx {
  int64_t n_u = 0, n_p = 0, N = 0, T = 0, V = 0;
  std::vector<int64_t> rowptr, pm_rowptr;
  std::vector<int32_t> col, pm_col;
  std::vector<double> J, Mp, R, delta, sol_owned, sol, sol_old;
  std::vector<double> xy;
  std::vector<int32_t> cv, cd;
  std::vector<int32_t> bf_cell, bf_face, bf_tag;
  Params prm{};
  // block-Jacobi extents for the ILU(0) preconditioners (one range per virtual rank)
  std::vector<int64_t> u_off{0}, p_off{0};
  std::vector<double> gmres_hist;  // residual estimate after every GMRES step of the last solve
  int64_t last_inner_its = 0;      // inner CG / GMRES iterations of the block preconditioner during the last solve
};

inline int64_t find_col(const Ctx &c, const std::vector<int64_t> &rp, const std::vector<int32_t> &cl,
                        int64_t row, int32_t colv) {
  (void)c;
  const int32_t *b = cl.data() + rp[row], *e = cl.data() + rp[row + 1];
  const int32_t *p = std::lower_bound(b, e, colv);
  if (p == e || *p != colv) return -1;
  return p - cl.data();
}

// ---------------------------------------------------------------------------------------
// assemble_system (cpp:178-378) / assemble_stokes_system (cpp:380-531): the literal loops.
// ---------------------------------------------------------------------------------------
void assemble(Ctx &c) {
  const Params &P = c.prm;
  std::fill(c.J.begin(), c.J.end(), 0.0);   // cpp:203
  std::fill(c.R.begin(), c.R.end(), 0.0);   // cpp:204
  std::fill(c.Mp.begin(), c.Mp.end(), 0.0); // cpp:205
  // reference-cell tables (FEValues construction, cpp:188-195)
  double psi[7][6], dpsi[7][6][2];
  for (int q = 0; q < 7; ++q) p2_eval(Q7.x[q], Q7.y[q], psi[q], dpsi[q]);
  const double gl[3] = {0.5 - 0.5 * std::sqrt(0.6), 0.5, 0.5 + 0.5 * std::sqrt(0.6)};
  const double gw[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};

  std::vector<std::vector<int32_t>> cell_bfaces;  // Neumann faces per cell
  std::vector<int32_t> nface_of(c.T, -1);
  for (size_t i = 0; i < c.bf_cell.size(); ++i)
    if (c.bf_tag[i] == P.neumann_id) {
      if (nface_of[c.bf_cell[i]] < 0) {
        nface_of[c.bf_cell[i]] = (int32_t)cell_bfaces.size();
        cell_bfaces.emplace_back();
      }
      cell_bfaces[nface_of[c.bf_cell[i]]].push_back(c.bf_face[i]);
    }

  double cell_matrix[15][15], cell_pm[15][15], cell_res[15];
  for (int64_t cell = 0; cell < c.T; ++cell) {
    const int32_t *v = &c.cv[3 * cell];
    const int32_t *dof = &c.cd[15 * cell];
    const double x0 = c.xy[2 * v[0]], y0 = c.xy[2 * v[0] + 1];
    const double J00 = c.xy[2 * v[1]] - x0, J01 = c.xy[2 * v[2]] - x0;
    const double J10 = c.xy[2 * v[1] + 1] - y0, J11 = c.xy[2 * v[2] + 1] - y0;
    const double det = J00 * J11 - J01 * J10;
    // J^{-T}
    const double a00 = J11 / det, a01 = -J10 / det, a10 = -J01 / det, a11 = J00 / det;
    for (int i = 0; i < 15; ++i) {
      cell_res[i] = 0;
      for (int j = 0; j < 15; ++j) cell_matrix[i][j] = cell_pm[i][j] = 0;
    }
    for (int q = 0; q < 7; ++q) {
      const double JxW = std::fabs(det) * Q7.w[q];
      // FEValues views of the 15 vector-valued shape functions at q
      double val[15][2], grad[15][2][2], dv[15], pv[15];
      for (int i = 0; i < 15; ++i) {
        int isp, comp, k;
        local_dof(i, isp, comp, k);
        val[i][0] = val[i][1] = 0;
        grad[i][0][0] = grad[i][0][1] = grad[i][1][0] = grad[i][1][1] = 0;
        dv[i] = 0;
        pv[i] = 0;
        if (isp) {
          const double lam[3] = {1 - Q7.x[q] - Q7.y[q], Q7.x[q], Q7.y[q]};
          pv[i] = lam[k];
        } else {
          const double gx = a00 * dpsi[q][k][0] + a01 * dpsi[q][k][1];
          const double gy = a10 * dpsi[q][k][0] + a11 * dpsi[q][k][1];
          val[i][comp] = psi[q][k];
          grad[i][comp][0] = gx;
          grad[i][comp][1] = gy;
          dv[i] = comp == 0 ? gx : gy;
        }
      }
      // get_function_values / gradients (cpp:229-233)
      double U[2] = {0, 0}, Uo[2] = {0, 0}, G[2][2] = {{0, 0}, {0, 0}}, Pq = 0;
      for (int i = 0; i < 15; ++i) {
        const double s = c.sol[dof[i]], so = c.sol_old[dof[i]];
        for (int a = 0; a < 2; ++a) {
          U[a] += s * val[i][a];
          Uo[a] += so * val[i][a];
          for (int b = 0; b < 2; ++b) G[a][b] += s * grad[i][a][b];
        }
        Pq += s * pv[i];
      }
      const double F[2] = {P.f[0], P.f[1]};  // forcing_term.vector_value (cpp:237-242)
      for (int i = 0; i < 15; ++i) {
        for (int j = 0; j < 15; ++j) {
          if (!P.stokes) {
            // Mass matrix (cpp:249-251)
            if (P.use_mass)
              cell_matrix[i][j] += (val[i][0] * val[j][0] + val[i][1] * val[j][1]) / P.deltat * JxW;
          }
          // Viscosity term (cpp:254-257 / 436-440)
          double sp = 0;
          for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) sp += grad[i][a][b] * grad[j][a][b];
          cell_matrix[i][j] += P.nu * P.rho * sp * JxW;
          if (!P.stokes) {
            // rho * (grad u^k * phi_j) * phi_i (cpp:259-263): Tensor<2>*Tensor<1> contracts the last index
            double t3 = 0;
            for (int a = 0; a < 2; ++a) {
              const double Gphi = G[a][0] * val[j][0] + G[a][1] * val[j][1];
              t3 += Gphi * val[i][a];
            }
            cell_matrix[i][j] += P.rho * t3 * JxW;
            // rho * (u^k * grad phi_j) * phi_i (cpp:265-269): Tensor<1>*Tensor<2> contracts u with the FIRST index
            double t4 = 0;
            for (int b = 0; b < 2; ++b) {
              const double Ug = U[0] * grad[j][0][b] + U[1] * grad[j][1][b];
              t4 += Ug * val[i][b];
            }
            cell_matrix[i][j] += P.rho * t4 * JxW;
          }
          // Pressure term in the momentum equation (cpp:272-274)
          cell_matrix[i][j] -= dv[i] * pv[j] * JxW;
          // Pressure term in the continuity equation (cpp:277-279)
          cell_matrix[i][j] -= dv[j] * pv[i] * JxW;
          // Pressure mass matrix (cpp:282-284)
          cell_pm[i][j] += pv[i] * pv[j] / P.nu * JxW;
        }
        if (!P.stokes) {
          // Time derivative term (cpp:288-290)
          if (P.use_mass)
            cell_res[i] -= P.rho * ((U[0] - Uo[0]) / P.deltat * val[i][0] + (U[1] - Uo[1]) / P.deltat * val[i][1]) * JxW;
          // viscous (cpp:292-295)
          double sp = 0;
          for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) sp += G[a][b] * grad[i][a][b];
          cell_res[i] -= P.nu * P.rho * sp * JxW;
          // convective (cpp:297-301): (u * grad u)_a = sum_c u_c G[c][a]
          double cv = 0;
          for (int a = 0; a < 2; ++a) cv += (U[0] * G[0][a] + U[1] * G[1][a]) * val[i][a];
          cell_res[i] -= P.rho * cv * JxW;
          // pressure (cpp:303-305)
          cell_res[i] += Pq * dv[i] * JxW;
        }
        // Forcing term (cpp:308-310 / 457-459)
        cell_res[i] += (F[0] * val[i][0] + F[1] * val[i][1]) * JxW;
      }
    }
    // Neumann boundary term (cpp:315-336 / 464-488)
    if (nface_of[cell] >= 0)
      for (int32_t f : cell_bfaces[nface_of[cell]]) {
        const int va = f, vb = (f + 1) % 3;
        const double ref[3][2] = {{0, 0}, {1, 0}, {0, 1}};
        const double ex = c.xy[2 * v[vb]] - c.xy[2 * v[va]], ey = c.xy[2 * v[vb] + 1] - c.xy[2 * v[va] + 1];
        const double L = std::sqrt(ex * ex + ey * ey);
        const double sgn = det > 0 ? 1.0 : -1.0;
        const double n[2] = {sgn * ey / L, -sgn * ex / L};
        for (int q = 0; q < 3; ++q) {
          const double s = gl[q];
          const double xr = ref[va][0] + s * (ref[vb][0] - ref[va][0]);
          const double yr = ref[va][1] + s * (ref[vb][1] - ref[va][1]);
          double ps[6], dps[6][2];
          p2_eval(xr, yr, ps, dps);
          const double JxW = L * gw[q];
          for (int i = 0; i < 15; ++i) {
            int isp, comp, k;
            local_dof(i, isp, comp, k);
            if (isp) continue;
            cell_res[i] += -P.p_out * (n[comp] * ps[k]) * JxW;
          }
        }
      }
    // global add (cpp:338-342); zeros are elided by Trilinos' add, a no-op numerically
    for (int i = 0; i < 15; ++i) {
      const int64_t r = dof[i];
      for (int j = 0; j < 15; ++j) {
        if (cell_matrix[i][j] != 0) {
          const int64_t p = find_col(c, c.rowptr, c.col, r, dof[j]);
          c.J[p] += cell_matrix[i][j];
        }
        if (cell_pm[i][j] != 0) {
          const int64_t p = find_col(c, c.pm_rowptr, c.pm_col, r, dof[j]);
          c.Mp[p] += cell_pm[i][j];
        }
      }
      c.R[r] += cell_res[i];
    }
  }
}

// MatrixTools::apply_boundary_values, Trilinos block version, eliminate_columns=false (cpp:375-376):
// per diagonal block, d = |first non-zero diagonal entry in the local row range|; constrained rows are cleared in
// every block of the block row (clear_rows(rows, d) on the diagonal block: the diagonal is ALWAYS set to d - the
// Trilinos version does not look at the old diagonal), solution_i = g_i, rhs_i = g_i * d.
// prm.dirichlet_diag == 1 gives deal.II's rule for its native SparseMatrix instead (a non-zero diagonal is kept,
// rhs_i = g_i * J_ii) - the two cannot be told apart without a deal.II run (recalled semantics).
// `sol` is delta_owned (Newton) or solution (Stokes).
void apply_dirichlet(Ctx &c, int64_t n, const int32_t *dofs, const double *vals, std::vector<double> &sol) {
  for (int block = 0; block < 2; ++block) {
    // "local row range" of a rank: with virtual ranks (orc_set_block_jacobi, the P > 1 comparisons) every rank takes d
    // from ITS rows of the block, as apply_boundary_values does on each MPI rank (d is rank-local)
    const std::vector<int64_t> &off = block == 0 ? c.u_off : c.p_off;
    const int64_t base = block == 0 ? 0 : c.n_u;
    for (size_t vr = 0; vr + 1 < off.size(); ++vr) {
      const int64_t r0 = base + off[vr], r1 = base + off[vr + 1];
      bool any = false;
      for (int64_t k = 0; k < n; ++k) any |= (dofs[k] >= r0 && dofs[k] < r1);
      if (!any) continue;
      double first_nz = 1;
      for (int64_t i = r0; i < r1; ++i) {
        const int64_t p = find_col(c, c.rowptr, c.col, i, (int32_t)i);
        if (p >= 0 && c.J[p] != 0) {
          first_nz = std::fabs(c.J[p]);
          break;
        }
      }
      for (int64_t k = 0; k < n; ++k) {
        const int64_t i = dofs[k];
        if (i < r0 || i >= r1) continue;
        const int64_t pd = find_col(c, c.rowptr, c.col, i, (int32_t)i);
        for (int64_t p = c.rowptr[i]; p < c.rowptr[i + 1]; ++p)
          if (p != pd) c.J[p] = 0;  // clears the row in the diagonal and the off-diagonal blocks
        double diag = first_nz;
        if (c.prm.dirichlet_diag == 1 && pd >= 0 && c.J[pd] != 0) diag = c.J[pd];
        if (pd >= 0) c.J[pd] = diag;
        sol[i] = vals[k];
        c.R[i] = vals[k] * diag;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// linear algebra on block vectors: a dot is the sum of the per-block dots (BlockVector)
// ---------------------------------------------------------------------------------------
using Vec = std::vector<double>;
struct Layout {
  int64_t n, split;  // split = n for a single-block vector
};
double dot(const Layout &L, const double *a, const double *b) {
  double s0 = 0, s1 = 0;
  for (int64_t i = 0; i < L.split; ++i) s0 += a[i] * b[i];
  for (int64_t i = L.split; i < L.n; ++i) s1 += a[i] * b[i];
  return s0 + s1;
}
double norm(const Layout &L, const double *a) { return std::sqrt(dot(L, a, a)); }

struct Csr {
  int64_t n = 0;
  std::vector<int64_t> rp;
  std::vector<int32_t> cl;
  Vec v;
  void vmult(double *y, const double *x) const {
    for (int64_t i = 0; i < n; ++i) {
      double s = 0;
      for (int64_t p = rp[i]; p < rp[i + 1]; ++p) s += v[p] * x[cl[p]];
      y[i] = s;
    }
  }
};

// BlockSparseMatrix::vmult: y_b = J_b0 x_0 + J_b1 x_1 (two separate row sums)
void jac_vmult(const Ctx &c, double *y, const double *x) {
  const int32_t nu = (int32_t)c.n_u;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < c.N; ++i) {
    double s0 = 0, s1 = 0;
    for (int64_t p = c.rowptr[i]; p < c.rowptr[i + 1]; ++p) {
      if (c.col[p] < nu)
        s0 += c.J[p] * x[c.col[p]];
      else
        s1 += c.J[p] * x[c.col[p]];
    }
    y[i] = s0 + s1;
  }
}

// rows [r0,r1) x cols [c0,c1) of a CSR, re-based to 0
Csr extract(const std::vector<int64_t> &rp, const std::vector<int32_t> &cl, const Vec &v, int64_t r0, int64_t r1,
            int64_t c0, int64_t c1) {
  Csr B;
  B.n = r1 - r0;
  B.rp.assign(B.n + 1, 0);
  for (int64_t i = r0; i < r1; ++i) {
    for (int64_t p = rp[i]; p < rp[i + 1]; ++p)
      if (cl[p] >= c0 && cl[p] < c1) {
        B.cl.push_back((int32_t)(cl[p] - c0));
        B.v.push_back(v[p]);
      }
    B.rp[i - r0 + 1] = (int64_t)B.cl.size();
  }
  return B;
}

// Ifpack_ILU, level-of-fill 0, atol 0, rtol 1, overlap 0 (TrilinosWrappers::PreconditionILU
// defaults, hpp:532-533): L unit lower, D = inverse pivots, U scaled by the inverse pivot.
// One independent factorisation per virtual rank's diagonal sub-block [off[r],off[r+1]).
struct Ilu0 {
  Csr F;                        // factored values on A's pattern restricted to the sub-blocks
  std::vector<int64_t> diag;    // position of the diagonal in each row
  Vec dinv;
  void compute(const Csr &A, const std::vector<int64_t> &off) {
    F.n = A.n;
    F.rp.assign(A.n + 1, 0);
    F.cl.clear();
    F.v.clear();
    for (size_t r = 0; r + 1 < off.size(); ++r)
      for (int64_t i = off[r]; i < off[r + 1]; ++i) {
        for (int64_t p = A.rp[i]; p < A.rp[i + 1]; ++p)
          if (A.cl[p] >= off[r] && A.cl[p] < off[r + 1]) {
            F.cl.push_back(A.cl[p]);
            F.v.push_back(A.v[p]);
          }
        F.rp[i + 1] = (int64_t)F.cl.size();
      }
    diag.assign(A.n, -1);
    dinv.assign(A.n, 0);
    std::vector<int64_t> colflag(A.n, -1);
    for (int64_t i = 0; i < A.n; ++i) {
      for (int64_t p = F.rp[i]; p < F.rp[i + 1]; ++p) {
        colflag[F.cl[p]] = p;
        if (F.cl[p] == i) diag[i] = p;
      }
      for (int64_t p = F.rp[i]; p < F.rp[i + 1] && F.cl[p] < i; ++p) {
        const int64_t j = F.cl[p];
        const double multiplier = F.v[p];
        F.v[p] *= dinv[j];
        for (int64_t q = diag[j] + 1; q < F.rp[j + 1]; ++q) {
          const int64_t kk = colflag[F.cl[q]];
          if (kk >= 0) F.v[kk] -= multiplier * F.v[q];
        }
      }
      dinv[i] = 1.0 / F.v[diag[i]];
      for (int64_t p = diag[i] + 1; p < F.rp[i + 1]; ++p) F.v[p] *= dinv[i];
      for (int64_t p = F.rp[i]; p < F.rp[i + 1]; ++p) colflag[F.cl[p]] = -1;
    }
  }
  void apply(double *y, const double *x) const {
    for (int64_t i = 0; i < F.n; ++i) {  // L solve, unit diagonal
      double s = x[i];
      for (int64_t p = F.rp[i]; p < diag[i]; ++p) s -= F.v[p] * y[F.cl[p]];
      y[i] = s;
    }
    for (int64_t i = 0; i < F.n; ++i) y[i] *= dinv[i];
    for (int64_t i = F.n - 1; i >= 0; --i) {  // U solve, unit diagonal
      double s = y[i];
      for (int64_t p = diag[i] + 1; p < F.rp[i + 1]; ++p) s -= F.v[p] * y[F.cl[p]];
      y[i] = s;
    }
  }
};

using Op = std::function<void(double *, const double *)>;
struct SolveResult {
  int its = 0;
  double res = 0;
  bool ok = false;
};

// deal.II SolverGMRES, default AdditionalData: 30 temporary vectors -> 28 inner
// steps, left preconditioning, modified Gram-Schmidt with the every-5th-step re-orthogonalisation
// test, Givens rotations, stop when the preconditioned residual estimate <= tol.
// `basis` persists across restarts (the preconditioner sees the recycled vector as its dst).
SolveResult gmres(const Layout &L, const Op &A, const Op &Pinv, double *x, const double *b, double tol, int max_steps,
                  int n_tmp, std::vector<double> *hist) {
  const int64_t n = L.n;
  std::vector<Vec> tmp(n_tmp, Vec(n, 0.0));
  const int m = n_tmp - 2;
  std::vector<double> H((size_t)(n_tmp) * (n_tmp - 1), 0.0), gamma(n_tmp), ci(n_tmp - 1), si(n_tmp - 1), h(n_tmp - 1);
  auto Hm = [&](int i, int j) -> double & { return H[(size_t)i * (n_tmp - 1) + j]; };
  SolveResult R;
  int accumulated = 0, dim = 0;
  bool re_orth = false;
  Vec &v = tmp[0], &p = tmp[n_tmp - 1];
  enum { ITER, SUCCESS, FAILURE } state = ITER;
  auto check = [&](int step, double val) {
    if (val <= tol) return SUCCESS;
    if (step >= max_steps || std::isnan(val)) return FAILURE;
    return ITER;
  };
  do {
    std::fill(h.begin(), h.end(), 0.0);
    A(p.data(), x);
    for (int64_t i = 0; i < n; ++i) p[i] = -p[i] + b[i];  // p.sadd(-1,1,b)
    Pinv(v.data(), p.data());
    double rho = norm(L, v.data());
    R.res = rho;
    state = check(accumulated, rho);
    if (state != ITER) break;
    gamma[0] = rho;
    const double inv = 1. / rho;
    for (int64_t i = 0; i < n; ++i) v[i] *= inv;
    for (int inner = 0; inner < m && state == ITER; ++inner) {
      ++accumulated;
      Vec &vv = tmp[inner + 1];
      A(p.data(), tmp[inner].data());
      Pinv(vv.data(), p.data());
      dim = inner + 1;
      // modified_gram_schmidt
      double norm_vv_start = 0;
      const bool consider = (!re_orth) && (inner % 5 == 4);
      if (consider) norm_vv_start = norm(L, vv.data());
      h[0] = dot(L, vv.data(), tmp[0].data());
      for (int i = 1; i < dim; ++i) {
        const double a = -h[i - 1];
        const Vec &V = tmp[i - 1];
        for (int64_t k = 0; k < n; ++k) vv[k] += a * V[k];
        h[i] = dot(L, vv.data(), tmp[i].data());
      }
      {
        const double a = -h[dim - 1];
        const Vec &V = tmp[dim - 1];
        for (int64_t k = 0; k < n; ++k) vv[k] += a * V[k];
      }
      double norm_vv = std::sqrt(dot(L, vv.data(), vv.data()));
      bool done_mgs = false;
      if (consider) {
        if (norm_vv > 10. * norm_vv_start * std::sqrt(std::numeric_limits<double>::epsilon()))
          done_mgs = true;
        else
          re_orth = true;
      }
      if (!done_mgs && re_orth) {
        double htmp = dot(L, vv.data(), tmp[0].data());
        h[0] += htmp;
        for (int i = 1; i < dim; ++i) {
          const Vec &V = tmp[i - 1];
          for (int64_t k = 0; k < n; ++k) vv[k] += -htmp * V[k];
          htmp = dot(L, vv.data(), tmp[i].data());
          h[i] += htmp;
        }
        const Vec &V = tmp[dim - 1];
        for (int64_t k = 0; k < n; ++k) vv[k] += -htmp * V[k];
        norm_vv = std::sqrt(dot(L, vv.data(), vv.data()));
      }
      const double s = norm_vv;
      h[inner + 1] = s;
      if (s != 0) {
        const double is = 1. / s;
        for (int64_t k = 0; k < n; ++k) vv[k] *= is;
      }
      // givens_rotation(h, gamma, ci, si, inner)
      for (int i = 0; i < inner; ++i) {
        const double sn = si[i], cs = ci[i], dummy = h[i];
        h[i] = cs * dummy + sn * h[i + 1];
        h[i + 1] = -sn * dummy + cs * h[i + 1];
      }
      const double r = 1. / std::sqrt(h[inner] * h[inner] + h[inner + 1] * h[inner + 1]);
      si[inner] = h[inner + 1] * r;
      ci[inner] = h[inner] * r;
      h[inner] = ci[inner] * h[inner] + si[inner] * h[inner + 1];
      gamma[inner + 1] = -si[inner] * gamma[inner];
      gamma[inner] *= ci[inner];
      for (int i = 0; i < dim; ++i) Hm(i, inner) = h[i];
      rho = std::fabs(gamma[dim]);
      R.res = rho;
      if (hist) hist->push_back(rho);
      state = check(accumulated, rho);
    }
    // H1.backward(h, gamma), then x += sum h_i v_i
    std::vector<double> y(dim);
    for (int i = dim - 1; i >= 0; --i) {
      double s = gamma[i];
      for (int j = i + 1; j < dim; ++j) s -= y[j] * Hm(i, j);
      y[i] = s / Hm(i, i);
    }
    for (int i = 0; i < dim; ++i) {
      const Vec &V = tmp[i];
      for (int64_t k = 0; k < n; ++k) x[k] += y[i] * V[k];
    }
  } while (state == ITER);
  R.its = accumulated;
  R.ok = state == SUCCESS;
  return R;
}

// deal.II SolverCG, preconditioned variant.
SolveResult cg(const Layout &L, const Op &A, const Op &Pinv, double *x, const double *b, double tol, int max_steps) {
  const int64_t n = L.n;
  Vec g(n), d(n), h(n);
  SolveResult R;
  bool all_zero = true;
  for (int64_t i = 0; i < n && all_zero; ++i) all_zero = x[i] == 0;
  if (!all_zero) {
    A(g.data(), x);
    for (int64_t i = 0; i < n; ++i) g[i] -= b[i];
  } else
    for (int64_t i = 0; i < n; ++i) g[i] = -b[i];
  double res = norm(L, g.data());
  R.res = res;
  if (res <= tol) {
    R.ok = true;
    return R;
  }
  if (0 >= max_steps || std::isnan(res)) return R;
  Pinv(h.data(), g.data());
  for (int64_t i = 0; i < n; ++i) d[i] = -h[i];
  double gh = dot(L, g.data(), h.data());
  int it = 0;
  while (true) {
    ++it;
    A(h.data(), d.data());
    double alpha = dot(L, d.data(), h.data());
    alpha = gh / alpha;
    for (int64_t i = 0; i < n; ++i) x[i] += alpha * d[i];
    for (int64_t i = 0; i < n; ++i) g[i] += alpha * h[i];
    res = std::sqrt(std::fabs(dot(L, g.data(), g.data())));
    R.res = res;
    R.its = it;
    if (res <= tol) {
      R.ok = true;
      return R;
    }
    if (it >= max_steps || std::isnan(res)) return R;
    Pinv(h.data(), g.data());
    double beta = gh;
    gh = dot(L, g.data(), h.data());
    beta = gh / beta;
    for (int64_t i = 0; i < n; ++i) d[i] = beta * d[i] - h[i];
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// C ABI for ctypes (mirrors include/nsg.h so the parity tests read symmetrically)
// ---------------------------------------------------------------------------------------
extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

void *orc_create(int64_t n_u, int64_t n_p, const int64_t *rowptr, const int32_t *col, const int64_t *pm_rowptr,
                 const int32_t *pm_col, int64_t T, int64_t V, const double *xy, const int32_t *cell_vertices,
                 const int32_t *cell_dofs, int64_t n_bf, const int32_t *bf_cell, const int32_t *bf_face,
                 const int32_t *bf_tag) {
  auto *c = new Ctx;
  c->n_u = n_u, c->n_p = n_p, c->N = n_u + n_p, c->T = T, c->V = V;
  c->rowptr.assign(rowptr, rowptr + c->N + 1);
  c->col.assign(col, col + rowptr[c->N]);
  c->pm_rowptr.assign(pm_rowptr, pm_rowptr + c->N + 1);
  c->pm_col.assign(pm_col, pm_col + pm_rowptr[c->N]);
  c->J.assign(c->col.size(), 0);
  c->Mp.assign(c->pm_col.size(), 0);
  for (Vec *v : {&c->R, &c->delta, &c->sol_owned, &c->sol, &c->sol_old}) v->assign(c->N, 0);
  c->xy.assign(xy, xy + 2 * V);
  c->cv.assign(cell_vertices, cell_vertices + 3 * T);
  c->cd.assign(cell_dofs, cell_dofs + 15 * T);
  c->bf_cell.assign(bf_cell, bf_cell + n_bf);
  c->bf_face.assign(bf_face, bf_face + n_bf);
  c->bf_tag.assign(bf_tag, bf_tag + n_bf);
  c->prm = Params{0.001, 1.0, 10.0, 0.05, {0, 0}, 10, 1, 0, 0};  // hpp:703-709, main.cpp:13, cpp:320
  c->u_off = {0, n_u};
  c->p_off = {0, n_p};
  return c;
}
void orc_destroy(void *h) { delete (Ctx *)h; }

void orc_set_params(void *h, double nu, double rho, double p_out, double deltat, double fx, double fy,
                    int32_t neumann_id, int32_t use_mass, int32_t stokes, int32_t dirichlet_diag) {
  ((Ctx *)h)->prm = Params{nu, rho, p_out, deltat, {fx, fy}, neumann_id, use_mass, stokes, dirichlet_diag};
}
void orc_set_block_jacobi(void *h, int n_parts, const int64_t *u_off, const int64_t *p_off) {
  Ctx *c = (Ctx *)h;
  c->u_off.assign(u_off, u_off + n_parts + 1);
  c->p_off.assign(p_off, p_off + n_parts + 1);
}
void orc_assemble(void *h) { assemble(*(Ctx *)h); }
void orc_apply_dirichlet(void *h, int64_t n, const int32_t *dofs, const double *vals, int32_t into_solution) {
  Ctx *c = (Ctx *)h;
  // into_solution: the values go to the ghosted `solution` (cpp:529), not to solution_owned
  apply_dirichlet(*c, n, dofs, vals, into_solution ? c->sol : c->delta);
}
double orc_residual_norm(void *h) {
  Ctx *c = (Ctx *)h;
  return norm(Layout{c->N, c->n_u}, c->R.data());
}
void orc_spmv(void *h, const double *x, double *y) { jac_vmult(*(Ctx *)h, y, x); }

// precond: 0 identity (cpp:570), 1 block-diagonal (hpp:520-572), 2 block-triangular (hpp:575-639).
// target: 0 -> delta_owned (solve_system), 1 -> solution_owned (solve_stokes_system).
// Returns 0 ok, 1 outer no-convergence, 2 inner solver no-convergence.
int orc_solve(void *h, int precond, double rel_tol, int max_it, int n_tmp, int target, int *its, double *res) {
  Ctx *c = (Ctx *)h;
  const Layout L{c->N, c->n_u};
  const double tol = rel_tol * norm(L, c->R.data());
  Op A = [c](double *y, const double *x) { jac_vmult(*c, y, x); };
  Csr Ab, Mb, Bb;
  Ilu0 ilu_a, ilu_m;
  bool inner_fail = false;
  int64_t inner_its = 0;  // iterations of the inner solvers, summed over the applications of the preconditioner
  Op Pinv;
  if (precond == 0) {
    Pinv = [c](double *y, const double *x) { std::memcpy(y, x, sizeof(double) * c->N); };
  } else {
    Ab = extract(c->rowptr, c->col, c->J, 0, c->n_u, 0, c->n_u);
    Mb = extract(c->pm_rowptr, c->pm_col, c->Mp, c->n_u, c->N, c->n_u, c->N);
    Bb = extract(c->rowptr, c->col, c->J, c->n_u, c->N, 0, c->n_u);
    ilu_a.compute(Ab, c->u_off);
    ilu_m.compute(Mb, c->p_off);
    const int64_t nu = c->n_u, np = c->n_p;
    Op Aop = [&Ab](double *y, const double *x) { Ab.vmult(y, x); };
    Op Mop = [&Mb](double *y, const double *x) { Mb.vmult(y, x); };
    Op Ia = [&ilu_a](double *y, const double *x) { ilu_a.apply(y, x); };
    Op Im = [&ilu_m](double *y, const double *x) { ilu_m.apply(y, x); };
    if (precond == 1)
      Pinv = [=, &inner_fail, &inner_its](double *y, const double *x) {
        const Layout Lu{nu, nu}, Lp{np, np};
        SolveResult r0 = gmres(Lu, Aop, Ia, y, x, 1e-2 * norm(Lu, x), 1000, 30, nullptr);
        SolveResult r1 = gmres(Lp, Mop, Im, y + nu, x + nu, 1e-2 * norm(Lp, x + nu), 1000, 30, nullptr);
        inner_its += r0.its + r1.its;
        if (!r0.ok || !r1.ok) inner_fail = true;
      };
    else
      Pinv = [=, &Bb, &inner_fail, &inner_its](double *y, const double *x) {
        const Layout Lu{nu, nu}, Lp{np, np};
        SolveResult r0 = cg(Lu, Aop, Ia, y, x, 1e-2 * norm(Lu, x), 2000);
        Vec tmp(np);
        Bb.vmult(tmp.data(), y);
        for (int64_t i = 0; i < np; ++i) tmp[i] = -tmp[i] + x[nu + i];  // tmp.sadd(-1, src1)
        SolveResult r1 = cg(Lp, Mop, Im, y + nu, tmp.data(), 1e-2 * norm(Lp, x + nu), 2000);
        inner_its += r0.its + r1.its;
        if (!r0.ok || !r1.ok) inner_fail = true;
      };
  }
  c->gmres_hist.clear();
  Vec &x = target ? c->sol_owned : c->delta;
  SolveResult r = gmres(L, A, Pinv, x.data(), c->R.data(), tol, max_it, n_tmp, &c->gmres_hist);
  if (its) *its = r.its;
  if (res) *res = r.res;
  c->last_inner_its = inner_its;
  if (target) c->sol = c->sol_owned;  // cpp:556
  if (inner_fail) return 2;
  return r.ok ? 0 : 1;
}
int64_t orc_last_inner_iterations(void *h) { return ((Ctx *)h)->last_inner_its; }
int64_t orc_gmres_history(void *h, double *out, int64_t cap) {
  Ctx *c = (Ctx *)h;
  const int64_t n = std::min<int64_t>(cap, (int64_t)c->gmres_hist.size());
  if (out) std::copy(c->gmres_hist.begin(), c->gmres_hist.begin() + n, out);
  return (int64_t)c->gmres_hist.size();
}
void orc_update_solution(void *h) {  // cpp:616-618
  Ctx *c = (Ctx *)h;
  for (int64_t i = 0; i < c->N; ++i) c->sol_owned[i] += c->delta[i];
  c->sol = c->sol_owned;
}
void orc_push_time_level(void *h) { ((Ctx *)h)->sol_old = ((Ctx *)h)->sol; }  // cpp:666
void orc_set_solution(void *h, const double *v) {
  Ctx *c = (Ctx *)h;
  c->sol_owned.assign(v, v + c->N);
  c->sol = c->sol_owned;
}
void orc_set_solution_old(void *h, const double *v) { ((Ctx *)h)->sol_old.assign(v, v + ((Ctx *)h)->N); }
void orc_set_delta(void *h, const double *v) { ((Ctx *)h)->delta.assign(v, v + ((Ctx *)h)->N); }
void orc_get_solution(void *h, double *o) { std::copy(((Ctx *)h)->sol.begin(), ((Ctx *)h)->sol.end(), o); }
void orc_get_delta(void *h, double *o) { std::copy(((Ctx *)h)->delta.begin(), ((Ctx *)h)->delta.end(), o); }
void orc_get_residual(void *h, double *o) { std::copy(((Ctx *)h)->R.begin(), ((Ctx *)h)->R.end(), o); }
void orc_get_matrix_values(void *h, double *o) { std::copy(((Ctx *)h)->J.begin(), ((Ctx *)h)->J.end(), o); }
void orc_get_pm_values(void *h, double *o) { std::copy(((Ctx *)h)->Mp.begin(), ((Ctx *)h)->Mp.end(), o); }

// N3 (not in the reference, which has no drag/lift code): force of the fluid on the body
// bounded by the faces with `boundary_id`,  F = -oint (rho nu grad(u) n - p n) ds  with n the outward
// normal of the FLUID domain, integrated with the reference's own 3-point face rule
// (QGaussSimplex<1>(3), cpp:52) and the reference's viscous form (grad u, not its symmetric part).
void orc_boundary_force(void *h, int32_t boundary_id, double *out2) {
  Ctx *c = (Ctx *)h;
  const double gl[3] = {0.5 - 0.5 * std::sqrt(0.6), 0.5, 0.5 + 0.5 * std::sqrt(0.6)};
  const double gw[3] = {5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0};
  const double ref[3][2] = {{0, 0}, {1, 0}, {0, 1}};
  double F[2] = {0, 0};
  for (size_t i = 0; i < c->bf_cell.size(); ++i) {
    if (c->bf_tag[i] != boundary_id) continue;
    const int64_t cell = c->bf_cell[i];
    const int f = c->bf_face[i];
    const int32_t *v = &c->cv[3 * cell];
    const int32_t *dof = &c->cd[15 * cell];
    const double x0 = c->xy[2 * v[0]], y0 = c->xy[2 * v[0] + 1];
    const double J00 = c->xy[2 * v[1]] - x0, J01 = c->xy[2 * v[2]] - x0;
    const double J10 = c->xy[2 * v[1] + 1] - y0, J11 = c->xy[2 * v[2] + 1] - y0;
    const double det = J00 * J11 - J01 * J10;
    const double a00 = J11 / det, a01 = -J10 / det, a10 = -J01 / det, a11 = J00 / det;
    const int va = f, vb = (f + 1) % 3;
    const double ex = c->xy[2 * v[vb]] - c->xy[2 * v[va]], ey = c->xy[2 * v[vb] + 1] - c->xy[2 * v[va] + 1];
    const double L = std::sqrt(ex * ex + ey * ey);
    const double sgn = det > 0 ? 1.0 : -1.0;
    const double n[2] = {sgn * ey / L, -sgn * ex / L};
    double fx = 0, fy = 0;
    for (int q = 0; q < 3; ++q) {
      const double s = gl[q];
      const double xr = ref[va][0] + s * (ref[vb][0] - ref[va][0]), yr = ref[va][1] + s * (ref[vb][1] - ref[va][1]);
      double ps[6], dps[6][2];
      p2_eval(xr, yr, ps, dps);
      double G[2][2] = {{0, 0}, {0, 0}};
      for (int k = 0; k < 6; ++k) {
        const double gx = a00 * dps[k][0] + a01 * dps[k][1], gy = a10 * dps[k][0] + a11 * dps[k][1];
        const int i0 = k < 3 ? 3 * k : 9 + 2 * (k - 3);
        for (int a = 0; a < 2; ++a) {
          G[a][0] += c->sol[dof[i0 + a]] * gx;
          G[a][1] += c->sol[dof[i0 + a]] * gy;
        }
      }
      const double lam[3] = {1 - xr - yr, xr, yr};
      double P = 0;
      for (int m = 0; m < 3; ++m) P += c->sol[dof[3 * m + 2]] * lam[m];
      const double w = L * gw[q];
      const double mu = c->prm.rho * c->prm.nu;
      fx += w * (mu * (G[0][0] * n[0] + G[0][1] * n[1]) - P * n[0]);
      fy += w * (mu * (G[1][0] * n[0] + G[1][1] * n[1]) - P * n[1]);
    }
    F[0] -= fx;
    F[1] -= fy;
  }
  out2[0] = F[0], out2[1] = F[1];
}

// ILU(0) apply on the velocity block, exposed for the K7 parity tests.
void orc_ilu_apply(void *h, int which, const double *x, double *y) {
  Ctx *c = (Ctx *)h;
  Csr B = which == 0 ? extract(c->rowptr, c->col, c->J, 0, c->n_u, 0, c->n_u)
                     : extract(c->pm_rowptr, c->pm_col, c->Mp, c->n_u, c->N, c->n_u, c->N);
  Ilu0 f;
  f.compute(B, which == 0 ? c->u_off : c->p_off);
  f.apply(y, x);
}

// Local 15x15 cell matrices of one triangle with prescribed nodal values, for the exact-integral
// known-answer tests: same code path as assemble() on a one-cell mesh is used by the tests, this
// only exports the quadrature rule so tests can check its degree of exactness.
void orc_quadrature(double *x, double *y, double *w) {
  for (int q = 0; q < 7; ++q) x[q] = Q7.x[q], y[q] = Q7.y[q], w[q] = Q7.w[q];
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
