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
#include "nsg_tri_layout.h"

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
// ---- one-CTA triangular solves ----------------------------------------------------------------------------------
// A 2-D P2 mesh in the reference's DoF order gives ~50 independent rows per dependency level and thousands of levels:
// the solve is one long chain, and what it costs is the latency of a dependency hop.  Through L2 (the stamped words
// above) a hop costs ~1.6 us.  Here ONE CTA walks the levels and a hop is a CTA barrier + a shared-memory load:
//  * rows are renumbered in level order ("positions"; inside a level by falling row length) and packed in groups of 4
//    consecutive positions = one warp, 8 lanes per row; lane c of a row owns its entries c, c + 8, c + 16, ...; a group's
//    entries are stored slab by slab (32 {factor, position} records, one per lane); every group of a level has the
//    level's slab count (short rows are padded with zero factors), so a group finds its records without a header;
//  * slabs and {right-hand side, inverse pivot} records stream through shared-memory rings: one producer thread hands
//    the TMA engine two bulk copies per level, TRI_DEPTH levels ahead, and waits on the level's mbarrier before it joins
//    the level barrier;
//  * the unknowns of the last TRI_W positions live in a shared-memory window and leave it through bulk copies,
//    TRI_FLUSH positions at a time; a column that has already left the window (the host stores ~position) is read back
//    from global memory.  Levels that do not fit the rings next to their TRI_DEPTH predecessors are read in place.
// The 8 lanes of a row hold the 32 partial sums a warp of the kernels above holds in its 32 lanes (4 each: entry j and
// j + 32 accumulate in "lane" j mod 32) and add them in the order of the xor-shuffle tree as lane 0 sees it (steps 16
// and 8 inside the thread, 4, 2, 1 by shuffles): the result is bitwise the same.  Measured on the 51 842-row block
// (1 120 + 1 120 levels), per apply: stamped solve 3.46 ms; a warp per row 3.67 ms, a thread per row 6.39 ms (one warp
// crawling through ~1 000 dependent instructions per level), 8 lanes per row with per-group headers and cp.async
// staging by 8 producer warps 2.0 ms (2 150 warp instructions per level on one SM), this kernel 1.67 ms = 0.75 us per
// level.  Tried on top of it and not kept (profiles/r02_summary.md): far unknowns prefetched into a shared-memory ring
// by the producer warp (2.09 - 2.11 ms), the next level's records prepared before the barrier with 24 compute warps
// (2.47 ms), rows and slots in one stream with a single bulk copy per level (3.04 ms).
constexpr int TRI_W = 8192, TRI_RQ = 8192, TRI_RK = 1024;
constexpr int TRI_CW = 20, TRI_MAX_THREADS = 1024;  // compute warps (the launch adds the producer warp; NSG_TRI_CW overrides);
                                                     // measured 6 / 8 / 12 / 16 / 20 / 24 / 31 warps: 2.33 / 2.12 / 1.87 / 1.87 / 1.70 / 1.73 / 1.75 ms
constexpr int TRI_DEPTH = 2;                                 // levels the producer runs ahead (a power of two). With 4, 407 of the 2 240
                                                             // levels of the 51 842-row block no longer fit the rings (tests/test_tri_layout.py)
constexpr int TRI_FLUSH = 512, TRI_MAX_LEVEL_ROWS = 2048;    // write-out granularity; widest level the window can take
constexpr size_t TRI_SMEM = sizeof(double) * TRI_W + 16 * (size_t)(TRI_RQ + TRI_RK);
struct TriArgs {
  int32_t n_levels;
  const int4 *info;     // [n_levels+1] {first slot, slabs per group, staged, first position}
  const int4 *slots;    // [slots] {factor (2 words), position of the column's unknown or ~position outside the window, 0}
  const double2 *rows;  // [n] by position: {right-hand side, inverse pivot}
  double *yg;           // [n] unknowns by position
  int32_t n;
};
__device__ __forceinline__ int4 lds128(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
  }
}
// one group of 4 rows (8 lanes each) of a level with S slabs per group; S = 5 stands for "5 or more" (n_slabs says how many)
template <bool UPPER, bool STAGED, int S>
__device__ __forceinline__ void tri_group(const TriArgs &A, uint32_t sm_y, uint32_t sm_slots, uint32_t sm_rows, int q0, int n_slabs, int k0,
                                          int nr, int lane) {
  const int sub = lane >> 3, k = k0 + sub;
  const bool writer = sub < nr && (lane & 7) == 0;
  double2 rr = make_double2(0.0, 1.0);
  if (writer) {
    if (STAGED) {
      const int4 w = lds128(sm_rows + 16u * (uint32_t)(k & (TRI_RK - 1)));
      rr = make_double2(__hiloint2double(w.y, w.x), __hiloint2double(w.w, w.z));
    } else {
      rr = A.rows[k];
    }
  }
  auto product_terms = [&](int i, double &f, double &yv) {
    const int q = q0 + 32 * i + lane;
    const int4 e = STAGED ? lds128(sm_slots + 16u * (uint32_t)(q & (TRI_RQ - 1))) : __ldg(A.slots + q);
    f = __hiloint2double(e.y, e.x);
    yv = lds64(sm_y + 8u * (uint32_t)(e.z & (TRI_W - 1)));
    if (e.z < 0) yv = __ldcg(A.yg + ~e.z);
  };
  double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < (S < 4 ? S : 4); ++i) {
    double f, yv;
    product_terms(i, f, yv);
    t[i] = __dmul_rn(f, yv);
  }
  if (S > 4) {  // rows with more than 32 entries: entry j + 32 accumulates on entry j
    for (int i0 = 4; i0 < n_slabs; i0 += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u < n_slabs) {
          double f, yv;
          product_terms(i0 + u, f, yv);
          t[u] = __fma_rn(f, yv, t[u]);
        }
    }
  }
  // the xor-16 and xor-8 steps of the tree; a step whose second operand is an untouched +0 is skipped (x + 0 = x)
  if (S > 2) t[0] = __dadd_rn(t[0], t[2]);
  if (S > 3) t[1] = __dadd_rn(t[1], t[3]);
  if (S > 1) t[0] = __dadd_rn(t[0], t[1]);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) t[0] = __dadd_rn(t[0], __shfl_xor_sync(0xffffffffu, t[0], o));
  if (writer) {
    const double v = UPPER ? __fma_rn(rr.x, rr.y, -t[0]) : __dsub_rn(rr.x, t[0]);
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(sm_y + 8u * (uint32_t)(k & (TRI_W - 1))), "d"(v) : "memory");
  }
}
template <bool UPPER, bool STAGED, int S>
__device__ __forceinline__ void tri_level(const TriArgs &A, uint32_t sm_y, uint32_t sm_slots, uint32_t sm_rows, const int4 lv, int warp, int lane,
                                          int n_warps) {
  const int n_slabs = lv.y & 0xffff, rows_in_level = lv.w - lv.z;
  for (int gi = warp; 4 * gi < rows_in_level; gi += n_warps)
    tri_group<UPPER, STAGED, S>(A, sm_y, sm_slots, sm_rows, lv.x + gi * n_slabs * 32, n_slabs, lv.z + 4 * gi, min(4, rows_in_level - 4 * gi), lane);
}
template <bool UPPER>
__global__ void __launch_bounds__(TRI_MAX_THREADS, 1) k_ilu_solve_cta(const TriArgs A) {
  extern __shared__ __align__(16) unsigned char tri_smem[];
  __shared__ int4 s_info[2 * TRI_DEPTH];  // {first slot, slabs | staged << 16, first position, one past the last position}
  __shared__ __align__(8) unsigned long long s_full[TRI_DEPTH];  // mbarriers: the copies of a level have landed
  const uint32_t sm_y = smem_addr_u32(tri_smem), sm_slots = sm_y + 8u * TRI_W, sm_rows = sm_slots + 16u * TRI_RQ;
  double *s_y = reinterpret_cast<double *>(tri_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cw = (int)(blockDim.x >> 5) - 1;  // compute warps
  const bool producer = tid == 32 * cw;       // lane 0 of the last warp
  if (producer) {
    for (int i = 0; i < TRI_DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(s_full + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto ring_copy = [&](uint32_t ring, int mask, const void *src, int first, int count, uint32_t bar) {  // 16-byte records
    const int s0 = first & mask, n1 = min(count, mask + 1 - s0);
    const char *g = reinterpret_cast<const char *>(src) + 16ll * first;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + 16u * (uint32_t)s0),
                 "l"(g), "r"(16 * n1), "r"(bar)
                 : "memory");
    if (n1 < count)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring), "l"(g + 16ll * n1),
                   "r"(16 * (count - n1)), "r"(bar)
                   : "memory");
  };
  auto stage_level = [&](int l) {  // producer: start the copies of level l and leave its extent for the compute warps
    const int4 a = __ldg(A.info + l), b = __ldg(A.info + l + 1);
    s_info[l & (2 * TRI_DEPTH - 1)] = make_int4(a.x, a.y | (a.z << 16), a.w, b.w);
    if (!a.z) return;
    const uint32_t bar = smem_addr_u32(s_full + (l & (TRI_DEPTH - 1)));
    const int n_slots = b.x - a.x, n_rows = b.w - a.w;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16 * (n_slots + n_rows)) : "memory");
    if (n_slots > 0) ring_copy(sm_slots, TRI_RQ - 1, A.slots, a.x, n_slots, bar);
    ring_copy(sm_rows, TRI_RK - 1, A.rows, a.w, n_rows, bar);
  };
  // Unknowns leave the window through the bulk-copy engine, TRI_FLUSH positions at a time (a plain global store per row
  // costs more instructions and has to drain at the level barriers).
  int flushed = 0;
  auto flush = [&](int upto) {  // producer; positions [flushed, upto) are final and visible (barrier), both even
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the previous write-out is complete: far readers may rely on it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int sa = flushed & (TRI_W - 1), cnt = upto - flushed, n1 = min(cnt, TRI_W - sa);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.yg + flushed), "r"(sm_y + 8u * (uint32_t)sa), "r"(8 * n1)
                 : "memory");
    if (n1 < cnt)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(A.yg + flushed + n1), "r"(sm_y), "r"(8 * (cnt - n1))
                   : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    flushed = upto;
  };
  uint32_t parities = 0;
  if (producer)
    for (int l = 0; l < TRI_DEPTH && l < A.n_levels; ++l) stage_level(l);
  for (int l = 0; l < A.n_levels; ++l) {
    if (producer && (s_info[l & (2 * TRI_DEPTH - 1)].y >> 16)) {  // the copies of level l have landed ...
      const int st = l & (TRI_DEPTH - 1);  // a level read in place does not use its mbarrier: count the phases per stage
      mbar_wait(smem_addr_u32(s_full + st), (parities >> st) & 1u);
      parities ^= 1u << st;
    }
    __syncthreads();  // ... and so have the unknowns of level l - 1; the rings of level l - 1 are free
    if (warp == cw) {
      if (producer) {
        const int final_below = s_info[l & (2 * TRI_DEPTH - 1)].z & ~1;  // the levels before l are complete
        if (final_below - flushed >= TRI_FLUSH) flush(final_below);
        if (l + TRI_DEPTH < A.n_levels) stage_level(l + TRI_DEPTH);
      }
    } else {
      const int4 lv = s_info[l & (2 * TRI_DEPTH - 1)];
      const int S = min(lv.y & 0xffff, 5);
#define NSG_TRI_CASE(n)                                                              \
  case n:                                                                            \
    if (lv.y >> 16) tri_level<UPPER, true, n>(A, sm_y, sm_slots, sm_rows, lv, warp, lane, cw); \
    else tri_level<UPPER, false, n>(A, sm_y, sm_slots, sm_rows, lv, warp, lane, cw);     \
    break;
      switch (S) {
        NSG_TRI_CASE(0)
        NSG_TRI_CASE(1)
        NSG_TRI_CASE(2)
        NSG_TRI_CASE(3)
        NSG_TRI_CASE(4)
        NSG_TRI_CASE(5)
      }
#undef NSG_TRI_CASE
    }
  }
  __syncthreads();
  if (producer) {
    flush((A.n + 1) & ~1);  // yg has room for the odd one
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  (void)s_y;
}
__global__ void k_scatter_by_index(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[idx[i]] = src[i];
}
// out[2 i] = vals[src[i]] (0 for padding): the factor word of the 16-byte slot records
__global__ void k_gather_slot_factors(int64_t n, const int64_t *__restrict__ src, const double *__restrict__ vals, double *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[2 * i] = src[i] >= 0 ? vals[src[i]] : 0.0;
}
// out[2 i + off] = src[idx[i]]: one word of the {right-hand side, inverse pivot} row records
__global__ void k_gather_row_word(int64_t n, const int32_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ out, int off) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[2 * i + off] = src[idx[i]];
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
  // level-order layouts of the one-CTA solves
  B.tri_ok = B.tri_pick = false;
  int32_t widest = 0;  // a level wider than TRI_MAX_LEVEL_ROWS would overrun the window before it is written out
  for (int32_t l = 0; l < nL; ++l) widest = std::max(widest, B.h_level_ptr[l + 1] - B.h_level_ptr[l]);
  for (int32_t l = 0; l < nU; ++l) widest = std::max(widest, B.h_ulevel_ptr[l + 1] - B.h_ulevel_ptr[l]);
  if (n > 0 && B.nnz < (int64_t)1 << 28 && widest <= TRI_MAX_LEVEL_ROWS) {
    double t_cta = 0.0;
    static_assert(sizeof(TriRec) == sizeof(int4), "TriRec is uploaded as int4");
    const TriLimits lim{TRI_W, TRI_RQ, TRI_RK, TRI_DEPTH};
    auto layout = [&](bool upper, const std::vector<int32_t> &rows_by_level, const std::vector<int32_t> &ptr,
                      const std::vector<int32_t> *pos_fwd, CsrBlock::Tri &T, std::vector<int32_t> &pos) -> int {
      TriLayout lay;  // nsg_tri_layout.h
      const int lrc = tri_layout(upper, n, rowptr.data(), col.data(), diag.data(), rows_by_level, ptr, pos_fwd, lim, lay);
      if (lrc) return fail(NSG_ERR_ARG, lrc == 1 ? "a row with more than 2^19 entries" : "the level-order layout of the triangular solve is too large");
      pos.swap(lay.pos);
      const int32_t nl = lay.n_levels;
      T.n_levels = nl, T.nq = lay.nq;
      // measured: ~0.35 us per staged level + ~0.4 us per pass of the compute warps over its groups (0.76 us per level on
      // the 51 842-row block), L2 round trips more per level read in place, 16 bytes per slot through one SM
      for (int32_t l = 0; l < nl; ++l) t_cta += 0.35e-6 + 0.4e-6 * ((ptr[l + 1] - ptr[l] + 4 * TRI_CW - 1) / (4 * TRI_CW));
      t_cta += 1.5e-6 * lay.in_place + 16.0 * (double)lay.nq / 60e9;
      NSG_TRY(upload(c, &T.info, reinterpret_cast<const int4 *>(lay.info.data()), nl + 1));
      NSG_TRY(upload(c, &T.slots, reinterpret_cast<const int4 *>(lay.slots.data()), (int64_t)lay.slots.size()));
      NSG_TRY(upload(c, &T.fsrc, lay.fsrc.data(), (int64_t)lay.fsrc.size()));
      NSG_TRY(upload(c, &T.ra_src, lay.ra_src.data(), n));
      NSG_TRY(upload(c, &T.rowid, lay.rows.data(), n));
      NSG_TRY(dev_alloc(&T.rows, n + 1));
      NSG_TRY(dev_alloc(&T.yg, n + 2));
      NSG_CUDA(cudaMemsetAsync(T.rows, 0, sizeof(double2) * (size_t)(n + 1), c->stream));
      NSG_CUDA(cudaStreamSynchronize(c->stream));  // the host vectors go out of scope
      return NSG_OK;
    };
    std::vector<int32_t> posL, posU;
    NSG_TRY(layout(false, rowsL, B.h_level_ptr, nullptr, B.triL, posL));
    NSG_TRY(layout(true, rowsU, B.h_ulevel_ptr, &posL, B.triU, posU));
    NSG_CUDA(cudaFuncSetAttribute(k_ilu_solve_cta<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_SMEM));
    NSG_CUDA(cudaFuncSetAttribute(k_ilu_solve_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRI_SMEM));
    B.tri_ok = true;
    B.tri_pick = t_cta < 1.55e-6 * (nL + nU);  // stamped single launch: ~1.55 us per level (profiles/r02_summary.md)
  }
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
  for (CsrBlock::Tri *T : {&B.triL, &B.triU})
    dev_free(T->info), dev_free(T->rowid), dev_free(T->slots), dev_free(T->fsrc), dev_free(T->ra_src), dev_free(T->rows), dev_free(T->yg);
  B.tri_ok = B.tri_pick = false;
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
  if (B.tri_ok) {  // the same factors in the level-order layouts of the one-CTA solves
    for (CsrBlock::Tri *T : {&B.triL, &B.triU})
      if (T->nq > 0) {
        k_gather_slot_factors<<<grid_for(T->nq, 256, 1 << 30), 256, 0, c->stream>>>(T->nq, T->fsrc, B.fval, reinterpret_cast<double *>(T->slots));
        NSG_LAUNCH_CHECK(c);
      }
    k_gather_row_word<<<grid_for(B.n, 256, 1 << 30), 256, 0, c->stream>>>(B.n, B.triU.rowid, B.dinv, reinterpret_cast<double *>(B.triU.rows), 1);
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
  const int variant = c->ilu_variant >= 0 ? c->ilu_variant : (B.tri_pick ? 2 : 1);
  if (variant == 2 && B.n > 0) {
    if (!B.tri_ok) return fail(NSG_ERR_ARG, "the one-CTA triangular solve cannot take this block (2^28 entries or a level of more than 2 048 rows)");
    const unsigned g = (unsigned)grid_for(B.n, 256, 1 << 30);
    static const int tri_cw = std::getenv("NSG_TRI_CW") ? std::min(31, std::max(1, std::atoi(std::getenv("NSG_TRI_CW")))) : TRI_CW;
    const unsigned tri_threads = 32u * (unsigned)(tri_cw + 1);
    auto args = [&](const CsrBlock::Tri &T) { return TriArgs{T.n_levels, T.info, T.slots, T.rows, T.yg, (int32_t)B.n}; };
    k_gather_row_word<<<g, 256, 0, c->stream>>>(B.n, B.triL.ra_src, x, reinterpret_cast<double *>(B.triL.rows), 0);
    NSG_LAUNCH_CHECK(c);
    k_ilu_solve_cta<false><<<1, tri_threads, TRI_SMEM, c->stream>>>(args(B.triL));
    NSG_LAUNCH_CHECK(c);
    k_gather_row_word<<<g, 256, 0, c->stream>>>(B.n, B.triU.ra_src, B.triL.yg, reinterpret_cast<double *>(B.triU.rows), 0);
    NSG_LAUNCH_CHECK(c);
    k_ilu_solve_cta<true><<<1, tri_threads, TRI_SMEM, c->stream>>>(args(B.triU));
    NSG_LAUNCH_CHECK(c);
    k_scatter_by_index<<<g, 256, 0, c->stream>>>(B.n, B.triU.rowid, B.triU.yg, y);
    NSG_LAUNCH_CHECK(c);
    return NSG_OK;
  }
  if (variant == 1 && B.n > 0) {
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
