// nsg_assemble.cuh — K1/K2/K3: owner-computes assembly of the Jacobian, pressure mass matrix and
// residual (reference: src/NavierStokesSolver.cpp:203-347), Neumann faces (cpp:315-336) and
// Dirichlet rows (cpp:349-377).
//
// Scheme ("row-owner"): every matrix row has exactly one owner thread.  A velocity P2 node owns
// rows (2n, 2n+1); a pressure vertex owns its Jacobian row and its pressure-mass row.  The owner
// walks the cells of its patch in ascending cell order (the order the reference's cell loop adds
// them), integrates only ITS rows of each 15x15 local matrix with the 7-point rule and adds them
// into a shared-memory image of the chunk's CSR rows.  The chunk image is then streamed to HBM
// once, fully coalesced: every Jacobian entry is written exactly once per assembly, there are no
// atomics, no zero-fill pass and the summation order is fixed.
#pragma once
#include "nsg_common.cuh"

namespace nsg {

constexpr int NPC = 128;  // row owners (threads) per CTA

// Reference-cell tables: QGaussSimplex<2>(3) (7 points, degree 5) and FE_SimplexP(2)/(1) values.
struct FeTables {
  double w[7];
  double psi[7][6];
  double dpsi[7][6][2];
  double chi[7][3];
  double mhat[6][6];  // sum_q w_q psi_k psi_l  (reference mass matrix of the same rule)
  double gl[3], gw[3];  // QGaussSimplex<1>(3) on [0,1]
};
__constant__ FeTables c_fe;

// Pre-integrated reference-cell tables for the factored kernels (variant 1).  Every entry is a sum over
// the SAME 7-point rule, so on an affine cell sum_q w_q f(q) factors exactly (up to rounding) into
// geometry x table: mass psi_k psi_l (also weighted by xi, eta for the affine velocity gradient),
// stiffness d_c psi_k d_d psi_l, divergence d_c psi_k chi_m, pressure mass chi_m chi_n, and the
// affine coefficients of the P2 reference gradients  d_c psi_l = ga + gb xi + gc eta.
struct FeTables2 {
  double Mh[6][6], Mx[6][6], My[6][6];
  double K00[6][6], K01s[6][6], K11[6][6];
  double Bh[6][3][2];
  double mh[6];
  double Mp[3][3];
  double ga[6][2], gb[6][2], gc[6][2];
  double qx[7], qy[7];
};
__constant__ FeTables2 c_fe2;

struct AsmParams {
  double nu, rho, p_out, dt_inv, f0, f1;
  int32_t use_mass, stokes, neumann_id;
  int32_t debug;  // NSG_ASM_DEBUG bit mask for bring-up experiments (0 in production)
};

// local scalar P2 index k -> position of its x-velocity dof in the 15-dof FESystem order
__host__ __device__ inline int uidx(int k) { return k < 3 ? 3 * k : 9 + 2 * (k - 3); }

__global__ void k_cell_geometry(int64_t T, const double *__restrict__ xy, const int32_t *__restrict__ cv,
                                double *__restrict__ geom) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= T) return;
  const int32_t v0 = cv[3 * c], v1 = cv[3 * c + 1], v2 = cv[3 * c + 2];
  const double x0 = xy[2 * v0], y0 = xy[2 * v0 + 1];
  const double J00 = xy[2 * v1] - x0, J01 = xy[2 * v2] - x0;
  const double J10 = xy[2 * v1 + 1] - y0, J11 = xy[2 * v2 + 1] - y0;
  const double det = J00 * J11 - J01 * J10;
  geom[5 * c + 0] = J11 / det;   // J^-T
  geom[5 * c + 1] = -J10 / det;
  geom[5 * c + 2] = -J01 / det;
  geom[5 * c + 3] = J00 / det;
  geom[5 * c + 4] = fabs(det);
}

// ---- velocity rows: A (both Frechet terms), B^T, residual --------------------------------------
#ifndef NSG_ASM_MINB
#define NSG_ASM_MINB 3
#endif
__global__ void __launch_bounds__(NPC, NSG_ASM_MINB)
k_assemble_u(const WorkList wl, const int64_t *__restrict__ rowptr, double *__restrict__ vals,
             double *__restrict__ R, const double *__restrict__ geom, const int32_t *__restrict__ cell_dofs,
             const double *__restrict__ sol, const double *__restrict__ sol_old, const AsmParams P) {
  extern __shared__ double s_vals[];
  __shared__ double s_psi[7][6], s_dpsi[7][6][2];  // only the entries indexed by the owner's (per-thread) k
  const int t = threadIdx.x;
  const int64_t b = blockIdx.x;
  const ChunkInfo ci = wl.chunks[b];
  const int ng = ci.g1 - ci.g0;
  const int64_t rs = rowptr[2 * (int64_t)ci.g0], re = rowptr[2 * (int64_t)ci.g1];
  const int cnt = (int)(re - rs);
  double *s_res = s_vals + cnt;
  for (int i = t; i < cnt + 2 * ng; i += NPC) s_vals[i] = 0.0;
  for (int i = t; i < 42; i += NPC) (&s_psi[0][0])[i] = (&c_fe.psi[0][0])[i];
  for (int i = t; i < 84; i += NPC) (&s_dpsi[0][0][0])[i] = (&c_fe.dpsi[0][0][0])[i];
  __syncthreads();

  const int desc = t < ci.n_threads ? wl.tdesc[b * NPC + t] : 0xffff;
  const bool have = desc != 0xffff;  // 0xffff: padding lane (an owner's slots never straddle a warp)
  const int gl = have ? (desc & 0xff) : 0, slot = have ? (desc >> 8) : 0;
  // commit rounds of THIS warp: the slots of an owner are adjacent lanes of one warp
  const int wrounds = __reduce_max_sync(0xffffffffu, have ? slot + 1 : 0);
  const int64_t node = ci.g0 + gl;
  const int64_t r0 = rowptr[2 * node];
  const int len = (int)(rowptr[2 * node + 1] - r0);
  double *row0 = s_vals + (r0 - rs), *row1 = row0 + len;
  const double nurho = P.nu * P.rho;

  for (int j = 0; j < ASM_PPT; ++j) {
    uint4 ra = make_uint4(0xffffffffu, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (have) {
      const uint4 *rp = reinterpret_cast<const uint4 *>(wl.recs + ci.rec_base + (int64_t)j * ci.n_threads + t);
      ra = __ldcs(rp);
      rb = __ldcs(rp + 1);
    }
    const bool work = (int)ra.x >= 0 && !(P.debug & 1);
    double A00[6], A01[6], A10[6], A11[6], B0[3], B1[3];
    double res0 = 0.0, res1 = 0.0;
#pragma unroll
    for (int l = 0; l < 6; ++l) A00[l] = A01[l] = A10[l] = A11[l] = 0.0;
#pragma unroll
    for (int m = 0; m < 3; ++m) B0[m] = B1[m] = 0.0;
    if (work) {
      const int64_t c = (int)ra.x;
      const int k = (int)ra.y;
      const double a00 = __ldg(geom + 5 * c), a01 = __ldg(geom + 5 * c + 1), a10 = __ldg(geom + 5 * c + 2),
                   a11 = __ldg(geom + 5 * c + 3), adet = __ldg(geom + 5 * c + 4);
      const int32_t *cd = cell_dofs + 15 * c;
      double u[6][2], pr[3];
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const int32_t d0 = __ldg(cd + uidx(l));
        u[l][0] = sol[d0];
        u[l][1] = sol[d0 + 1];
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) pr[m] = sol[__ldg(cd + 3 * m + 2)];
#pragma unroll
      for (int q = 0; q < 7; ++q) {  // fully unrolled: the reference-cell tables become constant-bank operands
        const double wq = adet * c_fe.w[q];
        double g[6][2];
        double U0 = 0, U1 = 0, G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const double dx = c_fe.dpsi[q][l][0], dy = c_fe.dpsi[q][l][1];
          g[l][0] = a00 * dx + a01 * dy;
          g[l][1] = a10 * dx + a11 * dy;
        }
        if (!P.stokes) {
#pragma unroll
          for (int l = 0; l < 6; ++l) {
            const double pl = c_fe.psi[q][l];
            U0 += u[l][0] * pl;
            U1 += u[l][1] * pl;
            G00 += u[l][0] * g[l][0];
            G01 += u[l][0] * g[l][1];
            G10 += u[l][1] * g[l][0];
            G11 += u[l][1] * g[l][1];
          }
        }
        const double pk = s_psi[q][k];
        const double gkx = a00 * s_dpsi[q][k][0] + a01 * s_dpsi[q][k][1];
        const double gky = a10 * s_dpsi[q][k][0] + a11 * s_dpsi[q][k][1];
        const double wpk = wq * pk;
        const double rw = P.rho * wpk;
        const double mk = (P.use_mass && !P.stokes) ? wpk * P.dt_inv : 0.0;
        const double vgx = nurho * wq * gkx, vgy = nurho * wq * gky;
        // A[(a,k),(b,l)] += w [ d_ab (psi_k psi_l/dt + nu rho g_k.g_l) + rho G_ab psi_k psi_l + rho U_b (g_l)_a psi_k ]
        const double d00 = mk + rw * G00, d01 = rw * G01, d10 = rw * G10, d11 = mk + rw * G11;
        const double c0 = rw * U0, c1 = rw * U1;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const double pl = c_fe.psi[q][l], glx = g[l][0], gly = g[l][1];
          const double visc = vgx * glx + vgy * gly;
          A00[l] += visc + d00 * pl + c0 * glx;
          A01[l] += d01 * pl + c1 * glx;
          A10[l] += d10 * pl + c0 * gly;
          A11[l] += visc + d11 * pl + c1 * gly;
        }
        // B^T[(a,k),m] -= w (g_k)_a chi_m   (cpp:272-274)
        const double bx = -wq * gkx, by = -wq * gky;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const double cm = c_fe.chi[q][m];
          B0[m] += bx * cm;
          B1[m] += by * cm;
        }
        // residual (cpp:287-311), time-derivative term added after the loop
        if (!P.stokes) {
          const double Pq = pr[0] * c_fe.chi[q][0] + pr[1] * c_fe.chi[q][1] + pr[2] * c_fe.chi[q][2];
          res0 += wq * (-nurho * (G00 * gkx + G01 * gky) - P.rho * (U0 * G00 + U1 * G10) * pk + Pq * gkx);
          res1 += wq * (-nurho * (G10 * gkx + G11 * gky) - P.rho * (U0 * G01 + U1 * G11) * pk + Pq * gky);
        }
        res0 += wpk * P.f0;
        res1 += wpk * P.f1;
      }
      if (!P.stokes && P.use_mass) {
        // -rho (u - u_old)/dt psi_k with the same 7-point rule = -rho/dt |detJ| sum_l mhat[k][l] (u_l - uold_l)
        double t0 = 0, t1 = 0;
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int32_t d0 = __ldg(cd + uidx(l));
          const double mh = c_fe.mhat[k][l];
          t0 += mh * (u[l][0] - sol_old[d0]);
          t1 += mh * (u[l][1] - sol_old[d0 + 1]);
        }
        const double f = -P.rho * P.dt_inv * adet;
        res0 += f * t0;
        res1 += f * t1;
      }
    }
    // commit rounds: slot r of every owner adds its pair into the owner's rows, in cell order
    const uint32_t ow[6] = {ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    for (int r = 0; r < wrounds; ++r) {
      if (work && slot == r && !(P.debug & 4)) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row0[o] += A00[l];
          row0[o + 1] += A01[l];
          row1[o] += A10[l];
          row1[o + 1] += A11[l];
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const int l = 6 + m;
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row0[o] += B0[m];
          row1[o] += B1[m];
        }
        s_res[2 * gl] += res0;
        s_res[2 * gl + 1] += res1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (!(P.debug & 2))
    for (int i = t; i < cnt; i += NPC) __stcs(vals + rs + i, s_vals[i]);
  for (int i = t; i < 2 * ng; i += NPC) R[2 * (int64_t)ci.g0 + i] = s_res[i];
}

// ---- pressure rows: B, the structurally present zero p-p block, pressure mass -------------------
__global__ void __launch_bounds__(NPC, 4)
k_assemble_p(const WorkList wl, int64_t n_own_u, const int64_t *__restrict__ rowptr, double *__restrict__ vals,
             const int64_t *__restrict__ pm_rowptr, double *__restrict__ pm_vals, double *__restrict__ R,
             const double *__restrict__ geom, const AsmParams P) {
  extern __shared__ double s_vals[];
  __shared__ double s_chi[7][3];  // indexed by the owner's (per-thread) m
  const int t = threadIdx.x;
  const int64_t b = blockIdx.x;
  const ChunkInfo ci = wl.chunks[b];
  const int ng = ci.g1 - ci.g0;
  const int64_t rs = rowptr[n_own_u + ci.g0], re = rowptr[n_own_u + ci.g1];
  const int64_t ms = pm_rowptr[n_own_u + ci.g0], me = pm_rowptr[n_own_u + ci.g1];
  const int cnt = (int)(re - rs), mcnt = (int)(me - ms);
  double *s_pm = s_vals + cnt;
  for (int i = t; i < cnt + mcnt; i += NPC) s_vals[i] = 0.0;
  for (int i = t; i < 21; i += NPC) (&s_chi[0][0])[i] = (&c_fe.chi[0][0])[i];
  __syncthreads();
  const int desc = t < ci.n_threads ? wl.tdesc[b * NPC + t] : 0xffff;
  const bool have = desc != 0xffff;  // 0xffff: padding lane (an owner's slots never straddle a warp)
  const int gl = have ? (desc & 0xff) : 0, slot = have ? (desc >> 8) : 0;
  // commit rounds of THIS warp: the slots of an owner are adjacent lanes of one warp
  const int wrounds = __reduce_max_sync(0xffffffffu, have ? slot + 1 : 0);
  const int64_t prow = n_own_u + ci.g0 + gl;
  double *row = s_vals + (rowptr[prow] - rs);
  double *mrow = s_pm + (pm_rowptr[prow] - ms);
  const double inv_nu = 1.0 / P.nu;
  for (int j = 0; j < ASM_PPT; ++j) {
    uint4 ra = make_uint4(0xffffffffu, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (have) {
      const uint4 *rp = reinterpret_cast<const uint4 *>(wl.recs + ci.rec_base + (int64_t)j * ci.n_threads + t);
      ra = __ldcs(rp);
      rb = __ldcs(rp + 1);
    }
    const bool work = (int)ra.x >= 0;
    double Bx[6], By[6], M[3];
#pragma unroll
    for (int l = 0; l < 6; ++l) Bx[l] = By[l] = 0.0;
    M[0] = M[1] = M[2] = 0.0;
    if (work) {
      const int64_t c = (int)ra.x;
      const int m = (int)ra.y;
      const double a00 = __ldg(geom + 5 * c), a01 = __ldg(geom + 5 * c + 1), a10 = __ldg(geom + 5 * c + 2),
                   a11 = __ldg(geom + 5 * c + 3), adet = __ldg(geom + 5 * c + 4);
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const double wq = adet * c_fe.w[q];
        const double cm = s_chi[q][m];
        const double wc = -wq * cm;
        // B[m,(b,l)] -= w (g_l)_b chi_m   (cpp:277-279)
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const double dx = c_fe.dpsi[q][l][0], dy = c_fe.dpsi[q][l][1];
          Bx[l] += wc * (a00 * dx + a01 * dy);
          By[l] += wc * (a10 * dx + a11 * dy);
        }
        // Mp[m,n] += w chi_m chi_n / nu   (cpp:282-284)
#pragma unroll
        for (int n = 0; n < 3; ++n) M[n] += cm * c_fe.chi[q][n] * inv_nu * wq;
      }
    }
    const uint32_t ow[6] = {ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    for (int r = 0; r < wrounds; ++r) {
      if (work && slot == r) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row[o] += Bx[l];
          row[o + 1] += By[l];
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          const int l = 6 + n;
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          mrow[o] += M[n];
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = t; i < cnt; i += NPC) __stcs(vals + rs + i, s_vals[i]);
  for (int i = t; i < mcnt; i += NPC) __stcs(pm_vals + ms + i, s_pm[i]);
  // no statement of the reference tests the pressure space: R_p == 0
  for (int i = t; i < ng; i += NPC) R[n_own_u + ci.g0 + i] = 0.0;
}

// ---- variant 1 ("factored"): the same integrals with the q-sum taken out where the integrand factors ----
// On an affine cell  g = A ghat (A = J^-T, d = |det J|):  mass and stiffness rows are geometry x table;
// grad u^k is affine in (xi,eta) so the first Frechet term is three table rows; only the second Frechet
// term and the convective residual keep a quadrature loop, and that loop runs in reference space (no
// per-point J^-T products).  ~0.45x the fp64 instructions of variant 0, same results to rounding.
#ifndef NSG_ASM2_MINB
#define NSG_ASM2_MINB 3
#endif
__global__ void __launch_bounds__(NPC, NSG_ASM2_MINB)
k_assemble_u2(const WorkList wl, const int64_t *__restrict__ rowptr, double *__restrict__ vals,
              double *__restrict__ R, const double *__restrict__ geom, const int32_t *__restrict__ cell_dofs,
              const double *__restrict__ sol, const double *__restrict__ sol_old, const AsmParams P) {
  extern __shared__ double s_vals[];
  // tables indexed by the owner's (per-thread) local index k live in shared memory
  __shared__ double s_psi[7][6], s_Mh[6][6], s_Mx[6][6], s_My[6][6], s_K00[6][6], s_K01s[6][6], s_K11[6][6], s_Bh[6][3][2],
      s_mh[6];
  const int t = threadIdx.x;
  const int64_t b = blockIdx.x;
  const ChunkInfo ci = wl.chunks[b];
  const int ng = ci.g1 - ci.g0;
  const int64_t rs = rowptr[2 * (int64_t)ci.g0], re = rowptr[2 * (int64_t)ci.g1];
  const int cnt = (int)(re - rs);
  double *s_res = s_vals + cnt;
  for (int i = t; i < cnt + 2 * ng; i += NPC) s_vals[i] = 0.0;
  for (int i = t; i < 42; i += NPC) (&s_psi[0][0])[i] = (&c_fe.psi[0][0])[i];
  for (int i = t; i < 36; i += NPC) {
    (&s_Mh[0][0])[i] = (&c_fe2.Mh[0][0])[i];
    (&s_Mx[0][0])[i] = (&c_fe2.Mx[0][0])[i];
    (&s_My[0][0])[i] = (&c_fe2.My[0][0])[i];
    (&s_K00[0][0])[i] = (&c_fe2.K00[0][0])[i];
    (&s_K01s[0][0])[i] = (&c_fe2.K01s[0][0])[i];
    (&s_K11[0][0])[i] = (&c_fe2.K11[0][0])[i];
    (&s_Bh[0][0][0])[i] = (&c_fe2.Bh[0][0][0])[i];
  }
  for (int i = t; i < 6; i += NPC) s_mh[i] = c_fe2.mh[i];
  __syncthreads();

  const int desc = t < ci.n_threads ? wl.tdesc[b * NPC + t] : 0xffff;
  const bool have = desc != 0xffff;  // 0xffff: padding lane (an owner's slots never straddle a warp)
  const int gl = have ? (desc & 0xff) : 0, slot = have ? (desc >> 8) : 0;
  // commit rounds of THIS warp: the slots of an owner are adjacent lanes of one warp
  const int wrounds = __reduce_max_sync(0xffffffffu, have ? slot + 1 : 0);
  const int64_t node = ci.g0 + gl;
  const int64_t r0 = rowptr[2 * node];
  const int len = (int)(rowptr[2 * node + 1] - r0);
  double *row0 = s_vals + (r0 - rs), *row1 = row0 + len;
  const double nurho = P.nu * P.rho;
  const bool ns = !P.stokes;

  for (int j = 0; j < ASM_PPT; ++j) {
    uint4 ra = make_uint4(0xffffffffu, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (have) {
      const uint4 *rp = reinterpret_cast<const uint4 *>(wl.recs + ci.rec_base + (int64_t)j * ci.n_threads + t);
      ra = __ldcs(rp);
      rb = __ldcs(rp + 1);
    }
    const bool work = (int)ra.x >= 0;
    double A00[6], A01[6], A10[6], A11[6], B0[3], B1[3];
    double res0 = 0.0, res1 = 0.0;
#pragma unroll
    for (int l = 0; l < 6; ++l) A00[l] = A01[l] = A10[l] = A11[l] = 0.0;
#pragma unroll
    for (int m = 0; m < 3; ++m) B0[m] = B1[m] = 0.0;
    if (work) {
      const int64_t c = (int)ra.x;
      const int k = (int)ra.y;
      const double a00 = __ldg(geom + 5 * c), a01 = __ldg(geom + 5 * c + 1), a10 = __ldg(geom + 5 * c + 2),
                   a11 = __ldg(geom + 5 * c + 3), d = __ldg(geom + 5 * c + 4);
      const int32_t *cd = cell_dofs + 15 * c;
      double u[6][2];
      double G0[2][2] = {{0, 0}, {0, 0}}, Gx[2][2] = {{0, 0}, {0, 0}}, Gy[2][2] = {{0, 0}, {0, 0}};
      double H[2][6][2];
      double cr0 = 0.0, cr1 = 0.0;
      if (ns) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int32_t d0 = __ldg(cd + uidx(l));
          u[l][0] = sol[d0];
          u[l][1] = sol[d0 + 1];
        }
        // reference gradient of u^k is affine: Ghat(xi,eta) = Gh0 + Ghx xi + Ghy eta; physical G_ab = sum_c A_bc Ghat_ac
        double h0[2][2] = {{0, 0}, {0, 0}}, hx[2][2] = {{0, 0}, {0, 0}}, hy[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
        for (int l = 0; l < 6; ++l)
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              h0[a][cc] += u[l][a] * c_fe2.ga[l][cc];
              hx[a][cc] += u[l][a] * c_fe2.gb[l][cc];
              hy[a][cc] += u[l][a] * c_fe2.gc[l][cc];
            }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          G0[a][0] = a00 * h0[a][0] + a01 * h0[a][1], G0[a][1] = a10 * h0[a][0] + a11 * h0[a][1];
          Gx[a][0] = a00 * hx[a][0] + a01 * hx[a][1], Gx[a][1] = a10 * hx[a][0] + a11 * hx[a][1];
          Gy[a][0] = a00 * hy[a][0] + a01 * hy[a][1], Gy[a][1] = a10 * hy[a][0] + a11 * hy[a][1];
        }
#pragma unroll
        for (int l = 0; l < 6; ++l) H[0][l][0] = H[0][l][1] = H[1][l][0] = H[1][l][1] = 0.0;
        // the only quadrature loop: H[b][l][c] = sum_q w psi_k U_b dhat_c psi_l, and the convective residual
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          double U0 = 0, U1 = 0;
#pragma unroll
          for (int l = 0; l < 6; ++l) {
            U0 += u[l][0] * c_fe.psi[q][l];
            U1 += u[l][1] * c_fe.psi[q][l];
          }
          const double wk = c_fe.w[q] * s_psi[q][k];
          const double c0 = wk * U0, c1 = wk * U1;
#pragma unroll
          for (int l = 0; l < 6; ++l) {
            H[0][l][0] += c0 * c_fe.dpsi[q][l][0];
            H[0][l][1] += c0 * c_fe.dpsi[q][l][1];
            H[1][l][0] += c1 * c_fe.dpsi[q][l][0];
            H[1][l][1] += c1 * c_fe.dpsi[q][l][1];
          }
          const double g00 = G0[0][0] + Gx[0][0] * c_fe2.qx[q] + Gy[0][0] * c_fe2.qy[q];
          const double g01 = G0[0][1] + Gx[0][1] * c_fe2.qx[q] + Gy[0][1] * c_fe2.qy[q];
          const double g10 = G0[1][0] + Gx[1][0] * c_fe2.qx[q] + Gy[1][0] * c_fe2.qy[q];
          const double g11 = G0[1][1] + Gx[1][1] * c_fe2.qx[q] + Gy[1][1] * c_fe2.qy[q];
          cr0 += wk * (U0 * g00 + U1 * g10);
          cr1 += wk * (U0 * g01 + U1 * g11);
        }
      }
      // S = A^T A: g_k . g_l = ghat_k^T S ghat_l
      const double S00 = a00 * a00 + a10 * a10, S01 = a00 * a01 + a10 * a11, S11 = a01 * a01 + a11 * a11;
      const double rd = P.rho * d, md = (P.use_mass && ns) ? P.dt_inv * d : 0.0, vd = nurho * d;
      double rv0 = 0.0, rv1 = 0.0;
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const double Mkl = s_Mh[k][l];
        const double Kkl = S00 * s_K00[k][l] + S01 * s_K01s[k][l] + S11 * s_K11[k][l];  // (1/d) sum_q w g_k.g_l
        const double D = md * Mkl + vd * Kkl;
        A00[l] = D;
        A11[l] = D;
        if (ns) {
          const double Mx = s_Mx[k][l], My = s_My[k][l];
          // rho w G_ab psi_k psi_l  (cpp:259-263)
          A00[l] += rd * (G0[0][0] * Mkl + Gx[0][0] * Mx + Gy[0][0] * My);
          A01[l] += rd * (G0[0][1] * Mkl + Gx[0][1] * Mx + Gy[0][1] * My);
          A10[l] += rd * (G0[1][0] * Mkl + Gx[1][0] * Mx + Gy[1][0] * My);
          A11[l] += rd * (G0[1][1] * Mkl + Gx[1][1] * Mx + Gy[1][1] * My);
          // rho w psi_k U_b (g_l)_a  (cpp:265-269): (g_l)_a = sum_c A_ac dhat_c psi_l
          A00[l] += rd * (a00 * H[0][l][0] + a01 * H[0][l][1]);
          A01[l] += rd * (a00 * H[1][l][0] + a01 * H[1][l][1]);
          A10[l] += rd * (a10 * H[0][l][0] + a11 * H[0][l][1]);
          A11[l] += rd * (a10 * H[1][l][0] + a11 * H[1][l][1]);
          rv0 += Kkl * u[l][0];
          rv1 += Kkl * u[l][1];
        }
      }
      // B^T[(a,k),m] = -sum_q w (g_k)_a chi_m = -d sum_c A_ac Bh[k][m][c]   (cpp:272-274)
      double pb0 = 0.0, pb1 = 0.0;
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const double bh0 = s_Bh[k][m][0], bh1 = s_Bh[k][m][1];
        const double t0 = a00 * bh0 + a01 * bh1, t1 = a10 * bh0 + a11 * bh1;
        B0[m] = -d * t0;
        B1[m] = -d * t1;
        if (ns) {
          const double pm = sol[__ldg(cd + 3 * m + 2)];
          pb0 += pm * t0;
          pb1 += pm * t1;
        }
      }
      // residual (cpp:287-311)
      if (ns) {
        res0 = -vd * rv0 - rd * cr0 + d * pb0;
        res1 = -vd * rv1 - rd * cr1 + d * pb1;
        if (P.use_mass) {
          double t0 = 0, t1 = 0;
#pragma unroll
          for (int l = 0; l < 6; ++l) {
            const int32_t d0 = __ldg(cd + uidx(l));
            const double mh = s_Mh[k][l];
            t0 += mh * (u[l][0] - sol_old[d0]);
            t1 += mh * (u[l][1] - sol_old[d0 + 1]);
          }
          res0 -= P.rho * P.dt_inv * d * t0;
          res1 -= P.rho * P.dt_inv * d * t1;
        }
      }
      res0 += P.f0 * d * s_mh[k];
      res1 += P.f1 * d * s_mh[k];
    }
    const uint32_t ow[6] = {ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    for (int r = 0; r < wrounds; ++r) {
      if (work && slot == r) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row0[o] += A00[l];
          row0[o + 1] += A01[l];
          row1[o] += A10[l];
          row1[o + 1] += A11[l];
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const int l = 6 + m;
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row0[o] += B0[m];
          row1[o] += B1[m];
        }
        s_res[2 * gl] += res0;
        s_res[2 * gl + 1] += res1;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = t; i < cnt; i += NPC) __stcs(vals + rs + i, s_vals[i]);
  for (int i = t; i < 2 * ng; i += NPC) R[2 * (int64_t)ci.g0 + i] = s_res[i];
}

// pressure rows, factored: B[m,(b,l)] = -d sum_c A_bc Bh[l][m][c];  Mp[m,n] = d/nu Mp_hat[m][n]
__global__ void __launch_bounds__(NPC, 4)
k_assemble_p2(const WorkList wl, int64_t n_own_u, const int64_t *__restrict__ rowptr, double *__restrict__ vals,
              const int64_t *__restrict__ pm_rowptr, double *__restrict__ pm_vals, double *__restrict__ R,
              const double *__restrict__ geom, const AsmParams P) {
  extern __shared__ double s_vals[];
  __shared__ double s_Bh[6][3][2], s_Mp[3][3];
  const int t = threadIdx.x;
  const int64_t b = blockIdx.x;
  const ChunkInfo ci = wl.chunks[b];
  const int ng = ci.g1 - ci.g0;
  const int64_t rs = rowptr[n_own_u + ci.g0], re = rowptr[n_own_u + ci.g1];
  const int64_t ms = pm_rowptr[n_own_u + ci.g0], me = pm_rowptr[n_own_u + ci.g1];
  const int cnt = (int)(re - rs), mcnt = (int)(me - ms);
  double *s_pm = s_vals + cnt;
  for (int i = t; i < cnt + mcnt; i += NPC) s_vals[i] = 0.0;
  for (int i = t; i < 36; i += NPC) (&s_Bh[0][0][0])[i] = (&c_fe2.Bh[0][0][0])[i];
  for (int i = t; i < 9; i += NPC) (&s_Mp[0][0])[i] = (&c_fe2.Mp[0][0])[i];
  __syncthreads();
  const int desc = t < ci.n_threads ? wl.tdesc[b * NPC + t] : 0xffff;
  const bool have = desc != 0xffff;  // 0xffff: padding lane (an owner's slots never straddle a warp)
  const int gl = have ? (desc & 0xff) : 0, slot = have ? (desc >> 8) : 0;
  // commit rounds of THIS warp: the slots of an owner are adjacent lanes of one warp
  const int wrounds = __reduce_max_sync(0xffffffffu, have ? slot + 1 : 0);
  const int64_t prow = n_own_u + ci.g0 + gl;
  double *row = s_vals + (rowptr[prow] - rs);
  double *mrow = s_pm + (pm_rowptr[prow] - ms);
  const double inv_nu = 1.0 / P.nu;
  for (int j = 0; j < ASM_PPT; ++j) {
    uint4 ra = make_uint4(0xffffffffu, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
    if (have) {
      const uint4 *rp = reinterpret_cast<const uint4 *>(wl.recs + ci.rec_base + (int64_t)j * ci.n_threads + t);
      ra = __ldcs(rp);
      rb = __ldcs(rp + 1);
    }
    const bool work = (int)ra.x >= 0;
    double Bx[6], By[6], M[3];
#pragma unroll
    for (int l = 0; l < 6; ++l) Bx[l] = By[l] = 0.0;
    M[0] = M[1] = M[2] = 0.0;
    if (work) {
      const int64_t c = (int)ra.x;
      const int m = (int)ra.y;
      const double a00 = __ldg(geom + 5 * c), a01 = __ldg(geom + 5 * c + 1), a10 = __ldg(geom + 5 * c + 2),
                   a11 = __ldg(geom + 5 * c + 3), d = __ldg(geom + 5 * c + 4);
#pragma unroll
      for (int l = 0; l < 6; ++l) {
        const double bh0 = s_Bh[l][m][0], bh1 = s_Bh[l][m][1];
        Bx[l] = -d * (a00 * bh0 + a01 * bh1);
        By[l] = -d * (a10 * bh0 + a11 * bh1);
      }
#pragma unroll
      for (int n = 0; n < 3; ++n) M[n] = s_Mp[m][n] * inv_nu * d;
    }
    const uint32_t ow[6] = {ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
    for (int r = 0; r < wrounds; ++r) {
      if (work && slot == r) {
#pragma unroll
        for (int l = 0; l < 6; ++l) {
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          row[o] += Bx[l];
          row[o + 1] += By[l];
        }
#pragma unroll
        for (int n = 0; n < 3; ++n) {
          const int l = 6 + n;
          const int o = (ow[l >> 1] >> ((l & 1) * 16)) & 0xffff;
          mrow[o] += M[n];
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = t; i < cnt; i += NPC) __stcs(vals + rs + i, s_vals[i]);
  for (int i = t; i < mcnt; i += NPC) __stcs(pm_vals + ms + i, s_pm[i]);
  for (int i = t; i < ng; i += NPC) R[n_own_u + ci.g0 + i] = 0.0;
}

// ---- K2: Neumann faces (cpp:315-336), one thread per boundary P2 node, faces in list order ------
__global__ void k_neumann(int64_t n_bnodes, const int32_t *__restrict__ bnode_dof, const int32_t *__restrict__ bnode_ptr,
                          const int32_t *__restrict__ bnode_face, const int32_t *__restrict__ bnode_pos,
                          const int32_t *__restrict__ bface_cell, const int32_t *__restrict__ bface_face,
                          const int32_t *__restrict__ bface_tag, const int32_t *__restrict__ cv,
                          const double *__restrict__ xy, double *__restrict__ R, const AsmParams P) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_bnodes) return;
  double r0 = 0, r1 = 0;
  bool any = false;
  for (int32_t e = bnode_ptr[i]; e < bnode_ptr[i + 1]; ++e) {
    const int32_t bf = bnode_face[e];
    if (bface_tag[bf] != P.neumann_id) continue;
    any = true;
    const int64_t c = bface_cell[bf];
    const int f = bface_face[bf];
    const int32_t va = cv[3 * c + f], vb = cv[3 * c + (f + 1) % 3], vc = cv[3 * c + (f + 2) % 3];
    const double ex = xy[2 * vb] - xy[2 * va], ey = xy[2 * vb + 1] - xy[2 * va + 1];
    const double L = sqrt(ex * ex + ey * ey);
    // outward normal: pointing away from the opposite vertex
    double nx = ey / L, ny = -ex / L;
    if (nx * (xy[2 * vc] - xy[2 * va]) + ny * (xy[2 * vc + 1] - xy[2 * va + 1]) > 0) nx = -nx, ny = -ny;
    const int pos = bnode_pos[e];  // 0: first vertex of the face, 1: second vertex, 2: midpoint
    double acc = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double s = c_fe.gl[q];
      const double ps = pos == 0 ? (1 - s) * (1 - 2 * s) : (pos == 1 ? s * (2 * s - 1) : 4 * s * (1 - s));
      acc += ps * (L * c_fe.gw[q]);
    }
    r0 += -P.p_out * nx * acc;
    r1 += -P.p_out * ny * acc;
  }
  if (any) {
    const int32_t d = bnode_dof[i];
    R[d] += r0;
    R[d + 1] += r1;
  }
}

// ---- N3: force of the fluid on the body bounded by the faces with `boundary_id` --------
// F = -oint (rho nu grad(u) n - p n) ds, n = outward normal of the fluid, 3-point face rule. One thread per
// boundary face writes its contribution (faces whose edge-midpoint node this rank owns); the per-face
// values are then summed in index order by one block (deterministic).
__global__ void k_face_force(int64_t n_bfaces, int32_t boundary_id, int64_t n_own_u, const int32_t *__restrict__ bface_cell,
                             const int32_t *__restrict__ bface_face, const int32_t *__restrict__ bface_tag,
                             const int32_t *__restrict__ cv, const int32_t *__restrict__ cell_dofs, const double *__restrict__ xy,
                             const double *__restrict__ geom, const double *__restrict__ sol, double mu, double *__restrict__ fx,
                             double *__restrict__ fy) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_bfaces) return;
  double ax = 0.0, ay = 0.0;
  const int64_t c = bface_cell[i];
  const int f = bface_face[i];
  const int32_t *cd = cell_dofs + 15 * c;
  if (bface_tag[i] == boundary_id && cd[uidx(3 + f)] < n_own_u) {
    const int32_t va = cv[3 * c + f], vb = cv[3 * c + (f + 1) % 3], vc = cv[3 * c + (f + 2) % 3];
    const double ex = xy[2 * vb] - xy[2 * va], ey = xy[2 * vb + 1] - xy[2 * va + 1];
    const double L = sqrt(ex * ex + ey * ey);
    double nx = ey / L, ny = -ex / L;
    if (nx * (xy[2 * vc] - xy[2 * va]) + ny * (xy[2 * vc + 1] - xy[2 * va + 1]) > 0) nx = -nx, ny = -ny;
    const double a00 = geom[5 * c], a01 = geom[5 * c + 1], a10 = geom[5 * c + 2], a11 = geom[5 * c + 3];
    const double rx[3] = {0, 1, 0}, ry[3] = {0, 0, 1};
    const int ia = f, ib = (f + 1) % 3;
    for (int q = 0; q < 3; ++q) {
      const double s = c_fe.gl[q];
      const double x = rx[ia] + s * (rx[ib] - rx[ia]), y = ry[ia] + s * (ry[ib] - ry[ia]);
      const double l0 = 1 - x - y, l1 = x, l2 = y;
      const double dp[6][2] = {{-(4 * l0 - 1), -(4 * l0 - 1)}, {4 * l1 - 1, 0.0}, {0.0, 4 * l2 - 1},
                               {4 * (l0 - l1), -4 * l1},       {4 * l2, 4 * l1},  {-4 * l2, 4 * (l0 - l2)}};
      double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
      for (int k = 0; k < 6; ++k) {
        const double gx = a00 * dp[k][0] + a01 * dp[k][1], gy = a10 * dp[k][0] + a11 * dp[k][1];
        const double u0 = sol[cd[uidx(k)]], u1 = sol[cd[uidx(k)] + 1];
        G00 += u0 * gx, G01 += u0 * gy, G10 += u1 * gx, G11 += u1 * gy;
      }
      const double P = sol[cd[2]] * l0 + sol[cd[5]] * l1 + sol[cd[8]] * l2;
      const double w = L * c_fe.gw[q];
      ax += w * (mu * (G00 * nx + G01 * ny) - P * nx);
      ay += w * (mu * (G10 * nx + G11 * ny) - P * ny);
    }
  }
  fx[i] = -ax;
  fy[i] = -ay;
}
__global__ void k_sum_ordered(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out2) {
  __shared__ double sa[256], sb[256];
  const int t = threadIdx.x;
  // contiguous slabs per thread, then an index-ordered tree: the same result for a given n
  const int64_t per = (n + 255) / 256, lo = min(n, t * per), hi = min(n, lo + per);
  double x = 0.0, y = 0.0;
  for (int64_t i = lo; i < hi; ++i) x += a[i], y += b[i];
  sa[t] = x, sb[t] = y;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) sa[t] += sa[t + s], sb[t] += sb[t + s];
    __syncthreads();
  }
  if (t == 0) out2[0] = sa[0], out2[1] = sb[0];
}

// ---- K3: MatrixTools::apply_boundary_values, Trilinos block version (cpp:375-376) ---
// position of the diagonal entry of every row (-1 if structurally absent), computed once
__global__ void k_diag_pos(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                           int64_t *__restrict__ diag_pos) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t lo = rowptr[i], hi = rowptr[i + 1], pd = -1;
  while (lo < hi) {  // columns ascend
    const int64_t mid = (lo + hi) >> 1;
    const int32_t cm = col[mid];
    if (cm == i) {
      pd = mid;
      break;
    }
    if (cm < i) lo = mid + 1; else hi = mid;
  }
  diag_pos[i] = pd;
}
// first non-zero diagonal entry of a diagonal block in the local range [r0,r1): index by atomicMin
// (the minimum is unique, so the result does not depend on the execution order), value afterwards
__global__ void k_first_nonzero_diag_index(int64_t r0, int64_t r1, const int64_t *__restrict__ diag_pos,
                                           const double *__restrict__ vals, unsigned long long *first_idx) {
  const int64_t i = r0 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= r1 || (unsigned long long)i >= *first_idx) return;
  const int64_t p = diag_pos[i];
  if (p >= 0 && vals[p] != 0.0) atomicMin(first_idx, (unsigned long long)i);
}
__global__ void k_first_nonzero_diag_value(const int64_t *__restrict__ diag_pos, const double *__restrict__ vals,
                                           const unsigned long long *first_idx, double *out) {
  const unsigned long long i = *first_idx;
  *out = (i == ~0ull) ? 1.0 : fabs(vals[diag_pos[i]]);
}

// one warp per constrained row: clear the row in every block, keep a non-zero diagonal (else the
// block's first non-zero diagonal), set the solution entry and rhs_i = g_i * diag_i
__global__ void k_apply_dirichlet(int64_t n, const int32_t *__restrict__ dofs, const double *__restrict__ g,
                                  int64_t n_own_u, const int64_t *__restrict__ rowptr, const int64_t *__restrict__ diag_pos,
                                  double *__restrict__ vals, double *__restrict__ x, double *__restrict__ R,
                                  const double *__restrict__ first_nz) {
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n) return;
  const int64_t i = dofs[w];
  const int64_t pd = diag_pos[i];
  for (int64_t p = rowptr[i] + lane; p < rowptr[i + 1]; p += 32)
    if (p != pd) vals[p] = 0.0;
  if (lane == 0) {
    double diag = pd >= 0 ? vals[pd] : 0.0;
    if (pd >= 0 && diag == 0.0) {
      diag = first_nz[i < n_own_u ? 0 : 1];
      vals[pd] = diag;
    }
    if (x) x[i] = g[w];
    R[i] = g[w] * diag;
  }
}

}  // namespace nsg
