// nsg_precond.cuh — K7: Ifpack-style ILU(0) with level-scheduled triangular solves, and the two block
// preconditioners of the reference (src/NavierStokesSolver.hpp:520-572 block-diagonal, :575-639
// block-triangular).  Included by nsg.cu after nsg_krylov.cuh.
//
//  * The Krylov operators (velocity block A, pressure mass Mp, B) are filtered SpMVs on the parent
//    CSR (rows of one block, columns of one block incl. its ghosts) — the distributed matrix blocks.
//  * TrilinosWrappers::PreconditionILU = Ifpack ILU(0), overlap 0: factor the rank-LOCAL diagonal
//    sub-block (owned rows x owned columns).  L unit lower, D = inverse pivots, U scaled by the
//    inverse pivot; apply = L-solve, D-scale, U-solve.  Rows of one dependency level
//    are processed in parallel, levels in order: the factor values equal the sequential ones.
#pragma once
#include "nsg_common.cuh"
#include "nsg_krylov.cuh"

namespace nsg {

// y[r] = sum over the columns of block `cb` (0: velocity incl. ghost u, 1: pressure incl. ghost p) of
// row r, rows [r0, r0+nr); 8 lanes per row.
__global__ void __launch_bounds__(256)
k_spmv_filtered(int64_t r0, int64_t nr, int cb, int64_t nu, int64_t nown, int64_t gu1, const int64_t *__restrict__ rowptr,
                const int32_t *__restrict__ col, const double *__restrict__ vals, const double *__restrict__ x,
                double *__restrict__ y, const int32_t *__restrict__ state) {
  if (state && *state != 0) return;
  const int64_t r = (blockIdx.x * (int64_t)256 + threadIdx.x) >> 3;
  const int l8 = threadIdx.x & 7;
  double acc = 0.0;
  if (r < nr) {
    const int64_t row = r0 + r;
    for (int64_t p = rowptr[row] + l8; p < rowptr[row + 1]; p += 8) {
      const int32_t cc = col[p];
      const bool is_u = cc < nu || (cc >= nown && cc < gu1);
      if (is_u == (cb == 0)) acc += vals[p] * x[cc];
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (r < nr && l8 == 0) y[r0 + r] = acc;
}

__global__ void k_gather_vals(int64_t n, const int64_t *__restrict__ src, const double *__restrict__ vals, double *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = vals[src[i]];
}

// one warp per row of the level: IKJ elimination restricted to the pattern (Ifpack_ILU::Compute)
__global__ void __launch_bounds__(128)
k_ilu_factor_level(const int32_t *__restrict__ rows, int nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const int64_t *__restrict__ diag, double *fval, double *dinv) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nrows) return;
  const int64_t i = rows[w];
  const int64_t s = rowptr[i], e = rowptr[i + 1], d = diag[i];
  for (int64_t p = s; p < d; ++p) {
    const int64_t j = col[p];
    const double multiplier = fval[p];
    __syncwarp();
    if (lane == 0) fval[p] = multiplier * dinv[j];
    for (int64_t q = diag[j] + 1 + lane; q < rowptr[j + 1]; q += 32) {
      const int32_t cq = col[q];
      int64_t lo = p + 1, hi = e;  // columns ascend and cq > j
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < cq) lo = mid + 1; else hi = mid;
      }
      if (lo < e && col[lo] == cq) fval[lo] -= multiplier * fval[q];
    }
    __syncwarp();
  }
  double di = 0.0;
  if (lane == 0) {
    di = 1.0 / fval[d];
    dinv[i] = di;
  }
  di = __shfl_sync(0xffffffffu, di, 0);
  for (int64_t p = d + 1 + lane; p < e; p += 32) fval[p] *= di;
}

// forward solve with unit lower factor, one warp per row of the level
__global__ void __launch_bounds__(128)
k_ilu_lsolve_level(const int32_t *__restrict__ rows, int nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const int64_t *__restrict__ diag, const double *__restrict__ fval, const double *__restrict__ x, double *y) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nrows) return;
  const int64_t i = rows[w];
  double acc = 0.0;
  for (int64_t p = rowptr[i] + lane; p < diag[i]; p += 32) acc += fval[p] * y[col[p]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[i] = x[i] - acc;
}
// D-scale folded into the backward solve with the unit upper factor
__global__ void __launch_bounds__(128)
k_ilu_usolve_level(const int32_t *__restrict__ rows, int nrows, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                   const int64_t *__restrict__ diag, const double *__restrict__ fval, const double *__restrict__ dinv, double *y) {
  const int w = (blockIdx.x * 128 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nrows) return;
  const int64_t i = rows[w];
  double acc = 0.0;
  for (int64_t p = diag[i] + 1 + lane; p < rowptr[i + 1]; p += 32) acc += fval[p] * y[col[p]];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[i] = y[i] * dinv[i] - acc;
}
// ---- single-launch triangular solves ------------------------------------------------------------------------
// The level-scheduled solves above need one launch per dependency level: thousands of ~3 us launches on a 2-D
// P2 mesh, each with a handful of rows.  Here ONE launch processes the rows in level order, one warp per row.
// A finished unknown is published as a 16-byte word {value, epoch stamp} written by ONE 128-bit store; a lane
// that needs y[j] polls that word with one 128-bit volatile load and gets value and stamp together: a
// dependency hop costs one L2 round trip and no fence (the packing trick of decoupled look-back scans).  CTAs
// take their position in the row order from a ticket counter, so a running CTA only ever waits for CTAs that
// started before it (resident or finished): no deadlock.  Lanes accumulate exactly as in the level kernels
// (stride 32, xor-shuffle tree): the result is bitwise the same.  Every wait is bounded (sf_error is raised
// instead of hanging the GPU).
constexpr int SF_WARPS = 4;
constexpr long long SF_TIMEOUT_CYCLES = 4000000000ll;  // ~2 s
__device__ __forceinline__ ulonglong2 ld_volatile_u128(const ulonglong2 *p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u128(ulonglong2 *p, unsigned long long a, unsigned long long b) {
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ double sf_wait_value(const ulonglong2 *word, unsigned long long epoch, int32_t *error, bool &ok) {
  ulonglong2 w = ld_volatile_u128(word);
  if (w.y != epoch) {
    const long long t0 = clock64();
    unsigned spins = 0;
    do {  // the poll itself is the only memory access on the dependency chain
      w = ld_volatile_u128(word);
      if (w.y == epoch) break;
      if ((++spins & 1023u) == 0u) {
        if (*(volatile int32_t *)error) {  // somebody already gave up: drain quickly
          ok = false;
          break;
        }
        if (clock64() - t0 > SF_TIMEOUT_CYCLES) {
          *(volatile int32_t *)error = 1;
          ok = false;
          break;
        }
      }
    } while (true);
  }
  return __longlong_as_double((long long)w.x);
}
// UPPER = false: yz[i] = {x[i] - sum_{j<i} L_ij y_j, epoch};  UPPER = true: reads the forward result from
// yz_in (complete: previous launch), yz[i] = {y_i dinv_i - sum_{j>i} U_ij y_j, epoch} and y[i] for the caller
template <bool UPPER>
__global__ void __launch_bounds__(32 * SF_WARPS)
k_ilu_solve_sf(const int32_t *__restrict__ rows, int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
               const int64_t *__restrict__ diag, const double *__restrict__ fval, const double *__restrict__ dinv,
               const double *__restrict__ x, const ulonglong2 *__restrict__ yz_in, ulonglong2 *yz, double *__restrict__ y,
               unsigned long long epoch, unsigned long long *ticket, unsigned long long ticket_base, int32_t *error) {
  // persistent CTAs: only a few hundred rows beyond the frontier are in flight, so the L2 is not flooded by
  // the polls of rows whose turn is far away (with one CTA per 4 rows the polls of ~10^4 resident warps
  // inflated every dependency hop from ~0.3 us to 2.6 us)
  __shared__ unsigned long long s_blk;
  const int lane = threadIdx.x & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_blk = atomicAdd(ticket, 1ull) - ticket_base;
    __syncthreads();
    const int64_t w = (int64_t)s_blk * SF_WARPS + (threadIdx.x >> 5);
    if ((int64_t)s_blk * SF_WARPS >= n) break;  // CTA-uniform
    if (w >= n) continue;
    const int64_t i = rows[w];
    const int64_t p0 = UPPER ? diag[i] + 1 : rowptr[i], p1 = UPPER ? rowptr[i + 1] : diag[i];
    double ra = 0.0, rb = 1.0;  // requested before the waits; combined below exactly as the level kernels do
    if (lane == 0) {
      if (UPPER) ra = __longlong_as_double((long long)yz_in[i].x), rb = dinv[i];
      else ra = x[i];
    }
    double acc = 0.0;
    bool ok = true;
    for (int64_t p = p0 + lane; p < p1; p += 32) {
      const double f = fval[p];
      acc += f * sf_wait_value(yz + col[p], epoch, error, ok);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const double v = UPPER ? ra * rb - acc : ra - acc;
      st_volatile_u128(yz + i, (unsigned long long)__double_as_longlong(v), epoch);
      if (UPPER) y[i] = v;
    }
    (void)ok;
  }
}
// y[off + i] = -y[off + i] + x[off + i]  (tmp.sadd(-1, src1), hpp:609)
__global__ void k_neg_add(int64_t n, double *__restrict__ y, const double *__restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = -y[i] + x[i];
}

// ---- host: local diagonal sub-block + level schedules (once per pattern) --------------------------
static int build_block(nsg_ctx *c, CsrBlock &B, const std::vector<int64_t> &rp, const std::vector<int32_t> &cl, int64_t r0, int64_t r1) {
  // rows [r0,r1) x columns [r0,r1) of the parent, re-based to 0: the rank-local matrix of Ifpack (overlap 0)
  const int64_t n = r1 - r0;
  std::vector<int64_t> rowptr(n + 1, 0), src, diag(n, -1);
  std::vector<int32_t> col;
  for (int64_t i = 0; i < n; ++i) {
    for (int64_t p = rp[r0 + i]; p < rp[r0 + i + 1]; ++p)
      if (cl[p] >= r0 && cl[p] < r1) {
        if (cl[p] - r0 == i) diag[i] = (int64_t)col.size();
        col.push_back((int32_t)(cl[p] - r0));
        src.push_back(p);
      }
    rowptr[i + 1] = (int64_t)col.size();
    if (diag[i] < 0) return fail(NSG_ERR_ARG, "a diagonal entry is missing from the sparsity pattern");
  }
  // dependency levels of the lower (forward) and upper (backward) solves
  std::vector<int32_t> levL(n, 0), levU(n, 0);
  int32_t nL = 0, nU = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t l = 0;
    for (int64_t p = rowptr[i]; p < diag[i]; ++p) l = std::max(l, levL[col[p]] + 1);
    levL[i] = l;
    nL = std::max(nL, l + 1);
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    int32_t l = 0;
    for (int64_t p = diag[i] + 1; p < rowptr[i + 1]; ++p) l = std::max(l, levU[col[p]] + 1);
    levU[i] = l;
    nU = std::max(nU, l + 1);
  }
  auto bucket = [&](const std::vector<int32_t> &lev, int32_t nl, std::vector<int32_t> &ptr, std::vector<int32_t> &rows) {
    ptr.assign(nl + 1, 0);
    for (int64_t i = 0; i < n; ++i) ptr[lev[i] + 1]++;
    for (int32_t l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
    rows.resize(n);
    std::vector<int32_t> pos(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[pos[lev[i]]++] = (int32_t)i;
  };
  std::vector<int32_t> rowsL, rowsU;
  bucket(levL, nL, B.h_level_ptr, rowsL);
  bucket(levU, nU, B.h_ulevel_ptr, rowsU);
  B.n = n;
  B.nnz = (int64_t)col.size();
  B.n_levels = nL;
  B.n_ulevels = nU;
  NSG_TRY(upload(c, &B.rowptr, rowptr.data(), n + 1));
  NSG_TRY(upload(c, &B.col, col.data(), B.nnz));
  NSG_TRY(upload(c, &B.src, src.data(), B.nnz));
  NSG_TRY(upload(c, &B.diag, diag.data(), n));
  NSG_TRY(upload(c, &B.level_rows, rowsL.data(), n));
  NSG_TRY(upload(c, &B.ulevel_rows, rowsU.data(), n));
  NSG_TRY(dev_alloc(&B.fval, B.nnz));
  NSG_TRY(dev_alloc(&B.dinv, n));
  NSG_TRY(dev_alloc(&B.done_l, n + 1));
  NSG_TRY(dev_alloc(&B.done_u, n + 1));
  NSG_TRY(dev_alloc(&B.ticket, 2));
  NSG_TRY(dev_alloc(&B.sf_error, 1));
  NSG_CUDA(cudaMemsetAsync(B.done_l, 0, sizeof(ulonglong2) * (size_t)(n + 1), c->stream));
  NSG_CUDA(cudaMemsetAsync(B.done_u, 0, sizeof(ulonglong2) * (size_t)(n + 1), c->stream));
  NSG_CUDA(cudaMemsetAsync(B.ticket, 0, 2 * sizeof(unsigned long long), c->stream));
  NSG_CUDA(cudaMemsetAsync(B.sf_error, 0, sizeof(int32_t), c->stream));
  B.epoch = 0, B.tickets_l = B.tickets_u = 0;
  NSG_CUDA(cudaStreamSynchronize(c->stream));
  return NSG_OK;
}

// Built lazily at the first use of a block preconditioner (the identity path never pays for it).
static int build_blocks(nsg_ctx *c) {
  if (c->have_blocks) return NSG_OK;
  {
    // the host copies of the patterns were released after nsg_set_mesh: read them back once
    std::vector<int64_t> rp(c->n_own + 1), prp(c->n_own + 1);
    std::vector<int32_t> cl(std::max<int64_t>(c->nnz, 1)), pcl(std::max<int64_t>(c->pm_nnz, 1));
    NSG_CUDA(cudaMemcpy(rp.data(), c->rowptr, 8 * rp.size(), cudaMemcpyDeviceToHost));
    NSG_CUDA(cudaMemcpy(cl.data(), c->col, 4 * (size_t)c->nnz, cudaMemcpyDeviceToHost));
    NSG_CUDA(cudaMemcpy(prp.data(), c->pm_rowptr, 8 * prp.size(), cudaMemcpyDeviceToHost));
    NSG_CUDA(cudaMemcpy(pcl.data(), c->pm_col, 4 * (size_t)c->pm_nnz, cudaMemcpyDeviceToHost));
    NSG_TRY(build_block(c, c->blkA, rp, cl, 0, c->n_own_u));
    NSG_TRY(build_block(c, c->blkM, prp, pcl, c->n_own_u, c->n_own));
  }
  NSG_TRY(dev_alloc(&c->inner_basis, 30 * c->stride));
  NSG_TRY(dev_alloc(&c->inner_ctl, 1));
  NSG_CUDA(cudaMallocHost((void **)&c->h_inner_ctl, sizeof(GmresCtl)));
  NSG_CUDA(cudaMemset(c->inner_ctl, 0, sizeof(GmresCtl)));
  c->have_blocks = true;
  c->blocks_stale = true;
  return NSG_OK;
}

static void free_block(CsrBlock &B) {
  dev_free(B.rowptr), dev_free(B.col), dev_free(B.src), dev_free(B.diag), dev_free(B.level_rows), dev_free(B.ulevel_rows);
  dev_free(B.fval), dev_free(B.dinv);
  dev_free(B.done_l), dev_free(B.done_u), dev_free(B.ticket), dev_free(B.sf_error);
}
static void free_blocks(nsg_ctx *c) {
  free_block(c->blkA);
  free_block(c->blkM);
  dev_free(c->inner_basis);
  dev_free(c->inner_ctl);
  if (c->h_inner_ctl) cudaFreeHost(c->h_inner_ctl);
  c->h_inner_ctl = nullptr;
}

static int ilu_factor(nsg_ctx *c, CsrBlock &B, const double *parent_vals) {
  if (B.nnz > 0) {
    k_gather_vals<<<grid_for(B.nnz, 256, 1 << 30), 256, 0, c->stream>>>(B.nnz, B.src, parent_vals, B.fval);
    NSG_LAUNCH_CHECK(c);
  }
  for (int32_t l = 0; l < B.n_levels; ++l) {
    const int nr = B.h_level_ptr[l + 1] - B.h_level_ptr[l];
    k_ilu_factor_level<<<(nr * 32 + 127) / 128, 128, 0, c->stream>>>(B.level_rows + B.h_level_ptr[l], nr, B.rowptr, B.col, B.diag,
                                                                     B.fval, B.dinv);
    NSG_LAUNCH_CHECK(c);
  }
  return NSG_OK;
}

// PreconditionBlock*::initialize (hpp:526-533, 582-590): ILU(0) of the current A and Mp blocks
static int precond_initialize(nsg_ctx *c) {
  NSG_TRY(build_blocks(c));
  if (!c->blocks_stale) return NSG_OK;
  NSG_TRY(ilu_factor(c, c->blkA, c->vals));
  NSG_TRY(ilu_factor(c, c->blkM, c->pm_vals));
  c->blocks_stale = false;
  return NSG_OK;
}

// y = (LU)^-1 x on block-local vectors of length B.n
static int ilu_apply(nsg_ctx *c, CsrBlock &B, double *y, const double *x) {
  if (c->ilu_variant == 1 && B.n > 0) {
    ++B.epoch;  // 64-bit stamp: never wraps
    // rows in flight = 4 x CTAs; measured on a 51 842-row block (1 120 levels): 148 CTAs 3.55 ms, 296: 3.64, 592: 3.69,
    // 1 184: 3.93, one CTA per 4 rows: 6.0 (level-scheduled launches: 9.3)
    static const int sf_grid = std::getenv("NSG_SF_GRID") ? std::max(1, std::atoi(std::getenv("NSG_SF_GRID"))) : sm_count();
    const unsigned grid = (unsigned)std::min<int64_t>((B.n + SF_WARPS - 1) / SF_WARPS, (int64_t)sf_grid);
    const unsigned long long used = (unsigned long long)((B.n + SF_WARPS - 1) / SF_WARPS) + grid;  // every CTA draws one ticket past the end
    k_ilu_solve_sf<false><<<grid, 32 * SF_WARPS, 0, c->stream>>>(B.level_rows, B.n, B.rowptr, B.col, B.diag, B.fval, B.dinv, x, nullptr,
                                                                 B.done_l, nullptr, B.epoch, B.ticket, B.tickets_l, B.sf_error);
    NSG_LAUNCH_CHECK(c);
    B.tickets_l += used;
    k_ilu_solve_sf<true><<<grid, 32 * SF_WARPS, 0, c->stream>>>(B.ulevel_rows, B.n, B.rowptr, B.col, B.diag, B.fval, B.dinv, nullptr, B.done_l,
                                                                B.done_u, y, B.epoch, B.ticket + 1, B.tickets_u, B.sf_error);
    NSG_LAUNCH_CHECK(c);
    B.tickets_u += used;
    return NSG_OK;
  }
  for (int32_t l = 0; l < B.n_levels; ++l) {
    const int nr = B.h_level_ptr[l + 1] - B.h_level_ptr[l];
    k_ilu_lsolve_level<<<(nr * 32 + 127) / 128, 128, 0, c->stream>>>(B.level_rows + B.h_level_ptr[l], nr, B.rowptr, B.col, B.diag,
                                                                     B.fval, x, y);
    NSG_LAUNCH_CHECK(c);
  }
  for (int32_t l = 0; l < B.n_ulevels; ++l) {
    const int nr = B.h_ulevel_ptr[l + 1] - B.h_ulevel_ptr[l];
    k_ilu_usolve_level<<<(nr * 32 + 127) / 128, 128, 0, c->stream>>>(B.ulevel_rows + B.h_ulevel_ptr[l], nr, B.rowptr, B.col, B.diag,
                                                                     B.fval, B.dinv, y);
    NSG_LAUNCH_CHECK(c);
  }
  return NSG_OK;
}

static int spmv_block(nsg_ctx *c, const int64_t *rowptr, const int32_t *col, const double *vals, int64_t r0, int64_t nr, int cb,
                      double *x_full, double *y_full, bool halo) {
  if (halo) NSG_TRY(halo_exchange(c, x_full));
  if (nr > 0) {
    k_spmv_filtered<<<(unsigned)((nr * 8 + 255) / 256), 256, 0, c->stream>>>(r0, nr, cb, c->n_own_u, c->n_own, c->n_own + c->n_ghost_u,
                                                                             rowptr, col, vals, x_full, y_full, nullptr);
    NSG_LAUNCH_CHECK(c);
  }
  return NSG_OK;
}

// dst = P^-1 src on full-layout vectors; dst keeps whatever it held as the inner solvers' initial guess
// (inside SolverGMRES that is a recycled temporary).
static int precond_vmult(nsg_ctx *c, int kind, double *dst, double *src) {
  const int64_t nu = c->n_own_u, np = c->n_own_p, S = c->stride;
  const Range ru{0, nu}, rpp{nu, np};
  double *wk = c->work + 4 * S;  // 4 vectors: 3 for CG + tmp (work[0..3] belong to the C-ABI test entry points)
  Op Aop = [c, nu](double *d, double *s, const int32_t *) { return spmv_block(c, c->rowptr, c->col, c->vals, 0, nu, 0, s, d, true); };
  Op Mop = [c, nu, np](double *d, double *s, const int32_t *) {
    return spmv_block(c, c->pm_rowptr, c->pm_col, c->pm_vals, nu, np, 1, s, d, true);
  };
  Op Ia = [c](double *d, double *s, const int32_t *) { return ilu_apply(c, c->blkA, d, s); };
  Op Im = [c, nu](double *d, double *s, const int32_t *) { return ilu_apply(c, c->blkM, d + nu, s + nu); };
  // tolerances 1e-2 * ||src_b|| (hpp:541-542, 550-551, 598-599, 611-612)
  NSG_TRY(dev_dot(c, nu, src, src, c->scal + 24, nullptr));
  NSG_TRY(dev_dot(c, np, src + nu, src + nu, c->scal + 25, nullptr));
  double n2[2];
  int32_t sf_err[2] = {0, 0};  // raised by the single-launch triangular solves of earlier applications
  NSG_CUDA(cudaMemcpyAsync(n2, c->scal + 24, 16, cudaMemcpyDeviceToHost, c->stream));
  if (c->blkA.sf_error) NSG_CUDA(cudaMemcpyAsync(&sf_err[0], c->blkA.sf_error, 4, cudaMemcpyDeviceToHost, c->stream));
  if (c->blkM.sf_error) NSG_CUDA(cudaMemcpyAsync(&sf_err[1], c->blkM.sf_error, 4, cudaMemcpyDeviceToHost, c->stream));
  NSG_CUDA(cudaStreamSynchronize(c->stream));
  c->d2h += 24;
  if (sf_err[0] || sf_err[1]) return fail(NSG_ERR_CUDA, "a single-launch triangular solve ran into its wait limit");
  const double tol_u = 1e-2 * std::sqrt(n2[0]), tol_p = 1e-2 * std::sqrt(n2[1]);
  if (kind == NSG_PRECOND_BLOCK_DIAGONAL) {
    GmresResult r0, r1;
    NSG_TRY(gmres_core(c, ru, Aop, &Ia, dst, src, src, 1e-2, 1000, 30, c->inner_basis, c->inner_ctl, c->h_inner_ctl, nullptr, 0, false, &r0));
    NSG_TRY(gmres_core(c, rpp, Mop, &Im, dst, src, src, 1e-2, 1000, 30, c->inner_basis, c->inner_ctl, c->h_inner_ctl, nullptr, 0, false, &r1));
    c->inner_its += r0.its + r1.its;
    if (!r0.ok || !r1.ok) return fail(NSG_ERR_INNER_NO_CONVERGENCE, "inner GMRES of the block-diagonal preconditioner did not converge");
  } else {
    CgResult r0, r1;
    NSG_TRY(cg_core(c, ru, Aop, Ia, dst, src, tol_u, 2000, wk, &r0));
    double *tmp = wk + 3 * S;
    // tmp = B dst0 ; tmp = -tmp + src1 (hpp:607-609)
    NSG_TRY(spmv_block(c, c->rowptr, c->col, c->vals, nu, np, 0, dst, tmp, true));
    k_neg_add<<<grid_for(np, 256), 256, 0, c->stream>>>(np, tmp + nu, src + nu);
    NSG_LAUNCH_CHECK(c);
    NSG_TRY(cg_core(c, rpp, Mop, Im, dst, tmp, tol_p, 2000, wk, &r1));
    c->inner_its += r0.its + r1.its;
    if (!r0.ok || !r1.ok) return fail(NSG_ERR_INNER_NO_CONVERGENCE, "inner CG of the block-triangular preconditioner did not converge");
  }
  return NSG_OK;
}

}  // namespace nsg
