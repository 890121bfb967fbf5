// nsg_gmres_fused.cuh — the whole identity-preconditioned SolverGMRES solve (src/NavierStokesSolver.cpp:566-583) as
// ONE cooperative kernel, for the meshes the reference actually ships (3e4 .. 3e5 DoFs).
//
// On those sizes a GMRES step is ~19 dependent kernels of 2-3 us each: 149 us per step with plain launches,
// 112 us as CUDA-graph segments (profiles/r01_summary.md) - the GPU idles between launches.  Here the grid stays
// resident for the whole solve (cudaLaunchCooperativeKernel), every thread keeps its entries of the vector being
// orthogonalised in REGISTERS across the modified Gram-Schmidt chain, a global inner product costs one stamped-slot
// exchange through L2 (no grid barrier), and the scalar bookkeeping (Givens rotations, stopping test, re-orthogonalisation test) never leaves
// the device.  The algorithm, its scalars (GmresCtl) and their update functions are those of gmres_core
// (nsg_krylov.cuh); only the partition of the inner-product sums differs, so the iterates agree with the
// multi-kernel path to rounding (not bitwise).  Used when: one rank, identity preconditioner, modified
// Gram-Schmidt, n <= GF_MAX_EPT x (resident threads); by default up to 65 536 unknowns.  Measured on B200
// (scripts/gmres_fused_check.py): 29 646 unknowns 111 -> 56 us per step; 117 324: 109 -> 101; 232 003: 118 -> 137 (one
// row per thread and one CTA per SM is too little memory parallelism there: the multi-kernel path stays the default).
#pragma once
#include <cooperative_groups.h>

#include "nsg_common.cuh"
#include "nsg_linalg.cuh"

namespace nsg {
constexpr int GF_MAX_GRID = 256;  // capacity of the partial-slot array: one CTA per SM, SM count taken from the device

namespace cg = cooperative_groups;

constexpr int GF_THREADS = 256;
constexpr int GF_MAX_EPT = 8;

__device__ __forceinline__ double gf_warp_allsum(double v) {  // every lane ends with the same bits
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum of one value per thread over the whole grid; every thread of every CTA returns identical bits (fixed
// partition, fixed order).  No grid barrier: a CTA publishes its partial as a 16-byte {value, epoch} word (one
// 128-bit store) and every warp polls the G words of the current epoch with 128-bit volatile loads - a global inner
// product costs one block reduction plus ~one L2 round trip (cg::grid.sync() measured ~4 us here).  Slots
// alternate between two halves; a half is reused two reductions later, when every CTA is provably past reading it
// (it had to publish the reduction in between).  Only scalars travel this way; whenever vector entries written by
// other threads are read next, a real grid barrier follows.
template <bool FENCE = false>
__device__ __forceinline__ double gf_grid_sum(double v, ulonglong2 *slots, unsigned long long &epoch, double *s_red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  ++epoch;
  ulonglong2 *half = slots + (epoch & 1ull) * gridDim.x;
  double *buf = s_red + (epoch & 1ull) * (GF_THREADS / 32 + 1);  // alternating: no barrier needed before the next reduction
  v = gf_warp_allsum(v);
  if (lane == 0) buf[wid] = v;
  __syncthreads();
  if (wid == 0) {  // ONE warp per CTA publishes and polls (8 polling warps per CTA made the L2 the bottleneck)
    double a = lane < GF_THREADS / 32 ? buf[lane] : 0.0;
    a = gf_warp_allsum(a);
    if (lane == 0) {
      // FENCE: the exchange doubles as a grid barrier for VECTOR data (release the CTA's stores, which the
      // __syncthreads above ordered before this thread; acquire below) - the scheme of cg::grid.sync()
      if (FENCE) __threadfence();
      unsigned long long bits = (unsigned long long)__double_as_longlong(a);
      asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(half + blockIdx.x), "l"(bits), "l"(epoch) : "memory");
    }
    double t = 0.0;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      ulonglong2 w;
      unsigned spins = 0;
      long long t0 = 0;
      do {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(half + i) : "memory");
        if (w.y == epoch) break;
        if ((++spins & 4095u) == 0u) {  // bounded: a lost CTA turns into NaN (solver failure), not into a hung GPU
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > 8000000000ll) {
            w.x = 0x7ff8000000000000ull;
            break;
          }
        }
      } while (true);
      t += __longlong_as_double((long long)w.x);
    }
    t = gf_warp_allsum(t);
    if (FENCE) __threadfence();
    if (lane == 0) buf[GF_THREADS / 32] = t;
  }
  __syncthreads();
  return buf[GF_THREADS / 32];
}

// {inv_s, state} of the scalar bookkeeping, computed by thread 0 of CTA 0, reach the other CTAs as one stamped 16-byte
// word (one L2 round trip) instead of through a grid barrier
__device__ __forceinline__ void gf_bcast_publish(ulonglong2 *word, unsigned long long epoch, double inv, int32_t state) {
  const unsigned long long lo = (unsigned long long)__double_as_longlong(inv), hi = (epoch << 16) | (unsigned long long)(state & 0xffff);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(word), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void gf_bcast_wait(const ulonglong2 *word, unsigned long long epoch, double *s_b, double &inv, int32_t &state) {
  if (threadIdx.x == 0) {
    ulonglong2 w;
    unsigned spins = 0;
    long long t0 = 0;
    do {
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(word) : "memory");
      if ((w.y >> 16) == epoch) break;
      if ((++spins & 4095u) == 0u) {
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 8000000000ll) {
          w.x = 0x7ff8000000000000ull, w.y = 2;  // state 2 = failure
          break;
        }
      }
    } while (true);
    s_b[0] = __longlong_as_double((long long)w.x);
    s_b[1] = (double)(int)(w.y & 0xffffull);
  }
  __syncthreads();
  inv = s_b[0];
  state = (int32_t)s_b[1];
  __syncthreads();
}

template <int EPT>
__global__ void __launch_bounds__(GF_THREADS)
k_gmres_solve_fused(int64_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ vals,
                    double *x, const double *__restrict__ b, double *basis, int64_t S, int n_tmp, GmresCtl *ctl, double *hist,
                    ulonglong2 *slots, unsigned long long *epoch_ctr) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_red[2 * (GF_THREADS / 32 + 1)];
  __shared__ double s_b[2];
  ulonglong2 *bword = slots + 2 * GF_MAX_GRID;  // broadcast word at a fixed place behind the (at most 2 x GF_MAX_GRID) partial slots
  const int64_t T = (int64_t)gridDim.x * GF_THREADS, tid = (int64_t)blockIdx.x * GF_THREADS + threadIdx.x;
  const int m = n_tmp - 2;
  // both counters continue where the previous solve stopped: stale stamps never match.  Reductions and broadcasts
  // count separately: the two halves of the partial slots must alternate strictly from one REDUCTION to the next.
  unsigned long long epoch = epoch_ctr[0], bepoch = epoch_ctr[1];
  bool re_orth = false;
  const double sqrt_eps = sqrt(2.220446049250313e-16);
  double vv[EPT];
  // row i of A times v, entries in CSR order (the order SpMV variant 0 uses)
  auto row_times = [&](int64_t i, const double *v) {
    double a = 0.0;
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) a += vals[p] * v[col[p]];
    return a;
  };
  auto V = [&](int j) { return basis + (int64_t)j * S; };

  while (true) {  // restart cycles
    if (*(volatile int32_t *)&ctl->state != 0) break;
    // ---- cycle start: p = b - A x ; v0 = p ; rho = ||v0|| ; v0 /= rho
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int64_t i = tid + k * T;
      vv[k] = 0.0;
      if (i < n) {
        vv[k] = -1.0 * row_times(i, x) + 1.0 * b[i];  // k_sadd(p, -1, 1, b)
        acc += vv[k] * vv[k];
      }
    }
    double tot = gf_grid_sum(acc, slots, epoch, s_red);
    ++bepoch;
    if (tid == 0) {
      ctl->nrm2 = tot;
      gm_cycle_start_dev(ctl);
      gf_bcast_publish(bword, bepoch, ctl->inv_s, ctl->state);
    }
    int32_t state;
    double inv;
    gf_bcast_wait(bword, bepoch, s_b, inv, state);
    if (state == 0) {
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int64_t i = tid + k * T;
        if (i < n) V(0)[i] = isfinite(inv) ? vv[k] * inv : vv[k];
      }
    }
    grid.sync();
    // ---- inner steps
    for (int inner = 0; inner < m && state == 0; ++inner) {
      const int dim = inner + 1;
      const double *vin = V(inner);
      double ns2 = 0.0;
      const bool consider = !re_orth && (inner % 5 == 4);
      acc = 0.0;
      double d0 = 0.0;
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int64_t i = tid + k * T;
        vv[k] = 0.0;
        if (i < n) {
          vv[k] = row_times(i, vin);
          acc += vv[k] * vv[k];
          d0 += vv[k] * V(0)[i];
        }
      }
      if (consider) ns2 = gf_grid_sum(acc, slots, epoch, s_red);
      // modified Gram-Schmidt: h_0 = vv.v_0 ; vv -= h_{j-1} v_{j-1}, h_j = vv.v_j ; ... ; nrm2 = vv.vv
      double hprev = gf_grid_sum(d0, slots, epoch, s_red);
      if (tid == 0) ctl->h[0] = hprev;
      for (int j = 1; j <= dim; ++j) {
        const double *vp = V(j - 1), *vn = V(j);
        acc = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int64_t i = tid + k * T;
          if (i < n) {
            vv[k] += (-1.0 * hprev) * vp[i];  // k_add_and_dot: vv += (sign * *aptr) * V
            acc += vv[k] * (j < dim ? vn[i] : vv[k]);
          }
        }
        hprev = gf_grid_sum(acc, slots, epoch, s_red);
        if (tid == 0) {
          if (j < dim) ctl->h[j] = hprev;
          else ctl->nrm2 = hprev;
        }
      }
      double nrm2 = hprev;
      if (consider) {
        if (tid == 0) ctl->norm_start2 = ns2;
        if (!(sqrt(nrm2) > 10. * sqrt(ns2) * sqrt_eps)) re_orth = true;
      }
      if (re_orth) {  // second sweep: the corrections go to h2 and are added in gm_step_dev
        d0 = 0.0;
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int64_t i = tid + k * T;
          if (i < n) d0 += vv[k] * V(0)[i];
        }
        hprev = gf_grid_sum(d0, slots, epoch, s_red);
        if (tid == 0) ctl->h2[0] = hprev;
        for (int j = 1; j <= dim; ++j) {
          const double *vp = V(j - 1), *vn = V(j);
          acc = 0.0;
#pragma unroll
          for (int k = 0; k < EPT; ++k) {
            const int64_t i = tid + k * T;
            if (i < n) {
              vv[k] += (-1.0 * hprev) * vp[i];
              acc += vv[k] * (j < dim ? vn[i] : vv[k]);
            }
          }
          hprev = gf_grid_sum(acc, slots, epoch, s_red);
          if (tid == 0) {
            if (j < dim) ctl->h2[j] = hprev;
            else ctl->nrm2 = hprev;
          }
        }
      }
      ++bepoch;
      if (tid == 0) {
        gm_step_dev(ctl, inner, re_orth ? 1 : 0, hist);
        gf_bcast_publish(bword, bepoch, ctl->inv_s, ctl->state);
      }
      gf_bcast_wait(bword, bepoch, s_b, inv, state);
      // the multi-kernel path scales only while state == 0 (k_scale_dev skips once the solver has decided)
      if (state == 0) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int64_t i = tid + k * T;
          if (i < n) V(inner + 1)[i] = isfinite(inv) ? vv[k] * inv : vv[k];
        }
      }
      (void)gf_grid_sum<true>(0.0, slots, epoch, s_red);  // grid barrier for the stored basis vector (next: SpMV gathers)
    }
    // ---- cycle end: x += sum_k y_k v_k
    if (tid == 0) {
      gm_backsolve_dev(ctl);
      __threadfence();
    }
    grid.sync();
    const int dim = *(volatile int32_t *)&ctl->dim;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int64_t i = tid + k * T;
      if (i < n) {
        double a = x[i];
        for (int q = 0; q < dim; ++q) a += *(volatile double *)&ctl->y[q] * V(q)[i];
        x[i] = a;
      }
    }
    grid.sync();
    if (state != 0) break;
  }
  if (tid == 0) epoch_ctr[0] = epoch, epoch_ctr[1] = bepoch;  // every thread counted the same exchanges
}

}  // namespace nsg
