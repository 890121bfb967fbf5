// nsg_krylov.cuh — host drivers of the Krylov solvers over the device kernels of nsg_linalg.cuh:
// deal.II SolverGMRES (outer solve_system, src/NavierStokesSolver.cpp:561-588, and the inner solves of
// PreconditionBlockDiagonal, hpp:543-558) and deal.II SolverCG (inner solves of
// PreconditionBlockTriangular, hpp:598-618).  Included by nsg.cu after its helper functions.
//
// All vectors are "full layout" device arrays [owned u | owned p | ghost u | ghost p]; a solver works
// on the sub-range [off, off+n) of them (the whole owned range for the outer solve, one block for the
// inner ones), so that an operator can refresh and read the ghosts of its input in place.
#pragma once
#include <functional>

#include "nsg_common.cuh"
#include "nsg_linalg.cuh"

namespace nsg {

struct Range {
  int64_t off, n;
};
// dst/src are full-layout base pointers; `state` (may be null) lets kernels skip once a solver has decided
using Op = std::function<int(double *dst, double *src, const int32_t *state)>;

__global__ void k_gmres_init(GmresCtl *c, double rel_tol, int max_steps, int n_tmp, int hist_cap) {
  c->tol = rel_tol * sqrt(c->nrm2);
  c->state = 0;
  c->accumulated = 0;
  c->dim = 0;
  c->max_steps = max_steps;
  c->n_tmp = n_tmp;
  c->hist_cap = hist_cap;
}

static int read_ctl_header(nsg_ctx *c, GmresCtl *dev, GmresCtl *host) {
  NSG_CUDA(cudaMemcpyAsync(host, dev, GM_HEADER_BYTES, cudaMemcpyDeviceToHost, c->stream));
  NSG_CUDA(cudaStreamSynchronize(c->stream));
  c->d2h += GM_HEADER_BYTES;
  return NSG_OK;
}

struct GmresResult {
  int its = 0;
  double res = 0;
  bool ok = false;
};

// deal.II SolverGMRES.  tol = rel_tol * ||tol_vec|| over the range.  `lazy`: no host
// synchronisation inside a restart cycle except at the every-5th-step re-orthogonalisation test
// (possible when neither operator needs the host); otherwise the header is read every step.
static int gmres_core(nsg_ctx *c, Range rg, const Op &A, const Op *Pinv, double *x, const double *b, const double *tol_vec,
                      double rel_tol, int max_steps, int n_tmp, double *basis, GmresCtl *ctl, GmresCtl *h_ctl, double *hist,
                      int hist_cap, bool lazy, GmresResult *out) {
  const int64_t n = rg.n, S = c->stride, o = rg.off;
  const int32_t *state = &ctl->state;
  auto V = [&](int i) { return basis + (int64_t)i * S; };
  double *p = V(n_tmp - 1);
  const int m = n_tmp - 2;
  const int vgrid = grid_for(n, 256);
  // temporaries start zeroed (a fresh TmpVectors pool); they are recycled across restarts
  NSG_CUDA(cudaMemsetAsync(basis, 0, 8 * (size_t)n_tmp * (size_t)S, c->stream));
  NSG_TRY(dev_dot(c, n, tol_vec + o, tol_vec + o, &ctl->nrm2, nullptr));
  k_gmres_init<<<1, 1, 0, c->stream>>>(ctl, rel_tol, max_steps, n_tmp, hist_cap);
  NSG_LAUNCH_CHECK(c);
  bool re_orth = false;
  const bool classical = c->orthogonalization == 1;

  // ---- the pieces of one restart cycle ----
  auto cycle_start = [&]() -> int {  // p = b - A x ; v0 = P^-1 p ; rho = ||v0|| ; v0 /= rho
    NSG_TRY(A(p, x, nullptr));
    k_sadd<<<vgrid, 256, 0, c->stream>>>(n, p + o, -1.0, 1.0, b + o);
    NSG_LAUNCH_CHECK(c);
    if (!Pinv)
      NSG_CUDA(cudaMemcpyAsync(V(0) + o, p + o, 8 * (size_t)n, cudaMemcpyDeviceToDevice, c->stream));
    else
      NSG_TRY((*Pinv)(V(0), p, nullptr));
    NSG_TRY(dev_dot(c, n, V(0) + o, V(0) + o, &ctl->nrm2, nullptr));
    k_gmres_cycle_start<<<1, 1, 0, c->stream>>>(ctl);
    NSG_LAUNCH_CHECK(c);
    k_scale_dev<<<vgrid, 256, 0, c->stream>>>(n, V(0) + o, &ctl->inv_s, state);
    NSG_LAUNCH_CHECK(c);
    return NSG_OK;
  };
  auto iter_front = [&](int inner, bool consider) -> int {  // A v, P^-1, first Gram-Schmidt sweep
    double *vv = V(inner + 1);
    const int dim = inner + 1;
    if (!Pinv) {
      NSG_TRY(A(vv, V(inner), state));
    } else {
      NSG_TRY(A(p, V(inner), state));
      NSG_TRY((*Pinv)(vv, p, state));
    }
    if (consider) NSG_TRY(dev_dot(c, n, vv + o, vv + o, &ctl->norm_start2, state));
    if (classical && dim <= CGS_MAXK) {  // h = V^T vv ; vv -= V h ; ||vv||  (two passes over the basis)
      NSG_TRY(dev_multi_dot(c, n, vv + o, basis + o, dim, &ctl->h[0], state));
      NSG_TRY(dev_multi_axpy_norm(c, n, vv + o, basis + o, &ctl->h[0], dim, &ctl->nrm2, state));
      return NSG_OK;
    }
    NSG_TRY(dev_dot(c, n, vv + o, V(0) + o, &ctl->h[0], state));
    for (int i = 1; i < dim; ++i)
      NSG_TRY(dev_add_and_dot(c, n, vv + o, &ctl->h[i - 1], -1.0, V(i - 1) + o, V(i) + o, &ctl->h[i], state));
    NSG_TRY(dev_add_and_dot(c, n, vv + o, &ctl->h[dim - 1], -1.0, V(dim - 1) + o, vv + o, &ctl->nrm2, state));
    return NSG_OK;
  };
  auto iter_back = [&](int inner, bool reorth_now) -> int {  // optional second sweep, Givens, scaling
    double *vv = V(inner + 1);
    const int dim = inner + 1;
    if (reorth_now && classical && dim <= CGS_MAXK) {
      NSG_TRY(dev_multi_dot(c, n, vv + o, basis + o, dim, &ctl->h2[0], state));
      NSG_TRY(dev_multi_axpy_norm(c, n, vv + o, basis + o, &ctl->h2[0], dim, &ctl->nrm2, state));
    } else if (reorth_now) {
      NSG_TRY(dev_dot(c, n, vv + o, V(0) + o, &ctl->h2[0], state));
      for (int i = 1; i < dim; ++i)
        NSG_TRY(dev_add_and_dot(c, n, vv + o, &ctl->h2[i - 1], -1.0, V(i - 1) + o, V(i) + o, &ctl->h2[i], state));
      NSG_TRY(dev_add_and_dot(c, n, vv + o, &ctl->h2[dim - 1], -1.0, V(dim - 1) + o, vv + o, &ctl->nrm2, state));
    }
    k_gmres_step<<<1, 1, 0, c->stream>>>(ctl, inner, reorth_now ? 1 : 0, hist);
    NSG_LAUNCH_CHECK(c);
    k_scale_dev<<<vgrid, 256, 0, c->stream>>>(n, vv + o, &ctl->inv_s, state);
    NSG_LAUNCH_CHECK(c);
    return NSG_OK;
  };
  auto cycle_end = [&]() -> int {  // x += sum_i y_i v_i with y from the rotated Hessenberg matrix
    k_gmres_backsolve<<<1, 1, 0, c->stream>>>(ctl);
    NSG_LAUNCH_CHECK(c);
    k_multi_axpy<<<vgrid, 256, 0, c->stream>>>(n, x + o, basis + o, S, ctl->y, &ctl->dim);
    NSG_LAUNCH_CHECK(c);
    return NSG_OK;
  };
  // after the consider-step test: true -> re-orthogonalise from now on
  auto reorth_test = [&]() {
    const double nv = std::sqrt(h_ctl->nrm2), ns = std::sqrt(h_ctl->norm_start2);
    return !(nv > 10. * ns * std::sqrt(std::numeric_limits<double>::epsilon()));
  };
  // run `body` either directly or as a cached graph (captured on first use)
  // graphs need launch segments made of kernels only: one rank, or all inter-rank traffic in peer-memory kernels
  // (fused all-reduce + peer-store halo; their sequence counters live on the device, so a replay stays in step)
  const bool use_graphs = lazy && c->use_graphs && (c->n_ranks == 1 || (c->halo_peer && c->peer.n_ranks > 1));
  auto run_segment = [&](int seg, const std::function<int()> &body) -> int {
    if (!use_graphs || seg < 0) return body();
    const GraphKey key{seg, n_tmp, n, o, x, b, basis, hist};
    for (GraphEntry &e : c->graphs)
      if (e.key == key) {
        NSG_CUDA(cudaGraphLaunch(e.exec, c->stream));
        c->launches += e.launches;
        c->last_solve[1]++;
        return NSG_OK;
      }
    const int64_t l0 = c->launches;
    NSG_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body();
    cudaGraph_t g = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
    if (rc != NSG_OK) return rc;
    if (ce != cudaSuccess) return fail(NSG_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    cudaGraphExec_t ex = nullptr;
    NSG_CUDA(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
    c->graphs.push_back(GraphEntry{key, ex, c->launches - l0});
    NSG_CUDA(cudaGraphLaunch(ex, c->stream));
    c->last_solve[1]++;
    return NSG_OK;
  };

  while (true) {
    if (lazy) {
      // segments end right after the first sweep of every 5th step (where the host must compare the norms
      // for the re-orthogonalisation test) and at the end of the cycle
      int inner = 0, seg = 0;
      bool pending = false;  // step `inner` has its first sweep done; Givens/scaling still to run
      while (true) {
        const int first = inner;
        const bool pb = pending, at_start = (first == 0 && !pending), ro = re_orth;
        int stop = pb ? first + 1 : first;                 // first step whose front this segment ends with
        while (stop < m && (ro || stop % 5 != 4)) ++stop;  // no test any more once re-orthogonalisation is on
        NSG_TRY(run_segment(ro ? -1 : seg, [&]() -> int {
          if (at_start) NSG_TRY(cycle_start());
          int i = first;
          if (pb) {
            NSG_TRY(iter_back(i, ro));
            ++i;
          }
          for (; i < stop; ++i) {
            NSG_TRY(iter_front(i, false));
            NSG_TRY(iter_back(i, ro));
          }
          if (stop < m)
            NSG_TRY(iter_front(stop, true));
          else
            NSG_TRY(cycle_end());
          return NSG_OK;
        }));
        ++seg;
        if (stop >= m) break;  // cycle complete, x updated
        NSG_TRY(read_ctl_header(c, ctl, h_ctl));
        if (h_ctl->state != 0) {  // decided in an earlier step: everything since was skipped
          NSG_TRY(cycle_end());
          break;
        }
        if (reorth_test()) re_orth = true;
        inner = stop;
        pending = true;
      }
    } else {
      NSG_TRY(cycle_start());
      NSG_TRY(read_ctl_header(c, ctl, h_ctl));
      if (h_ctl->state != 0) break;
      for (int inner = 0; inner < m; ++inner) {
        const bool consider = !re_orth && (inner % 5 == 4);
        NSG_TRY(iter_front(inner, consider));
        NSG_TRY(read_ctl_header(c, ctl, h_ctl));
        if (h_ctl->state != 0) break;
        if (consider && reorth_test()) re_orth = true;
        NSG_TRY(iter_back(inner, re_orth));
        NSG_TRY(read_ctl_header(c, ctl, h_ctl));
        if (h_ctl->state != 0) break;
      }
      NSG_TRY(cycle_end());
    }
    NSG_TRY(read_ctl_header(c, ctl, h_ctl));
    if (h_ctl->state != 0) break;
  }
  NSG_TRY(read_ctl_header(c, ctl, h_ctl));
  out->its = h_ctl->accumulated;
  out->res = h_ctl->rho;
  out->ok = (h_ctl->state & 0xff) == 1;
  return NSG_OK;
}

// scalar helpers for CG: device scalars s[0..] ; op 0: out = a / b ; 1: out = sqrt(|a|)
__global__ void k_scalar(int op, double *out, const double *a, const double *b) {
  if (op == 0)
    *out = *a / *b;
  else
    *out = sqrt(fabs(*a));
}
// d = -h  /  g = -b  /  g = g - b
__global__ void k_neg_copy(int64_t n, double *__restrict__ y, const double *__restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = -x[i];
}

struct CgResult {
  int its = 0;
  double res = 0;
  bool ok = false;
};

// deal.II SolverCG, preconditioned variant. tol is an absolute value known on the host.
// wk: three full-layout work vectors g, d, h. Device scalars in c->scal[16..24).
static int cg_core(nsg_ctx *c, Range rg, const Op &A, const Op &Pinv, double *x, const double *b, double tol, int max_steps,
                   double *wk, CgResult *out) {
  const int64_t n = rg.n, S = c->stride, o = rg.off;
  double *g = wk, *d = wk + S, *h = wk + 2 * S;
  double *s_gh = c->scal + 16, *s_dh = c->scal + 17, *s_alpha = c->scal + 18, *s_res2 = c->scal + 19, *s_beta = c->scal + 20,
         *s_x2 = c->scal + 21, *s_ghn = c->scal + 22;
  const int vgrid = grid_for(n, 256);
  double host2[2];
  auto read = [&](const double *dev, double *hst, int cnt) -> int {
    NSG_CUDA(cudaMemcpyAsync(hst, dev, 8 * cnt, cudaMemcpyDeviceToHost, c->stream));
    NSG_CUDA(cudaStreamSynchronize(c->stream));
    c->d2h += 8 * cnt;
    return NSG_OK;
  };
  // x.all_zero() short-circuit
  NSG_TRY(dev_dot(c, n, x + o, x + o, s_x2, nullptr));
  NSG_TRY(read(s_x2, host2, 1));
  if (host2[0] != 0.0) {
    NSG_TRY(A(g, x, nullptr));
    k_sadd<<<vgrid, 256, 0, c->stream>>>(n, g + o, 1.0, -1.0, b + o);
  } else {
    k_neg_copy<<<vgrid, 256, 0, c->stream>>>(n, g + o, b + o);
  }
  NSG_LAUNCH_CHECK(c);
  NSG_TRY(dev_dot(c, n, g + o, g + o, s_res2, nullptr));
  NSG_TRY(read(s_res2, host2, 1));
  double res = std::sqrt(host2[0]);
  out->res = res;
  out->its = 0;
  if (res <= tol) {
    out->ok = true;
    return NSG_OK;
  }
  if (0 >= max_steps || std::isnan(res)) return NSG_OK;
  NSG_TRY(Pinv(h, g, nullptr));
  k_neg_copy<<<vgrid, 256, 0, c->stream>>>(n, d + o, h + o);
  NSG_LAUNCH_CHECK(c);
  NSG_TRY(dev_dot(c, n, g + o, h + o, s_gh, nullptr));
  int it = 0;
  while (true) {
    ++it;
    NSG_TRY(A(h, d, nullptr));
    NSG_TRY(dev_dot(c, n, d + o, h + o, s_dh, nullptr));
    k_scalar<<<1, 1, 0, c->stream>>>(0, s_alpha, s_gh, s_dh);
    NSG_LAUNCH_CHECK(c);
    k_axpy_dev<<<vgrid, 256, 0, c->stream>>>(n, x + o, s_alpha, 1.0, d + o);
    NSG_LAUNCH_CHECK(c);
    NSG_TRY(dev_add_and_dot(c, n, g + o, s_alpha, 1.0, h + o, g + o, s_res2, nullptr));
    NSG_TRY(read(s_res2, host2, 1));
    res = std::sqrt(std::fabs(host2[0]));
    out->res = res;
    out->its = it;
    if (res <= tol) {
      out->ok = true;
      return NSG_OK;
    }
    if (it >= max_steps || std::isnan(res)) return NSG_OK;
    NSG_TRY(Pinv(h, g, nullptr));
    NSG_TRY(dev_dot(c, n, g + o, h + o, s_ghn, nullptr));
    k_scalar<<<1, 1, 0, c->stream>>>(0, s_beta, s_ghn, s_gh);
    NSG_LAUNCH_CHECK(c);
    NSG_CUDA(cudaMemcpyAsync(s_gh, s_ghn, 8, cudaMemcpyDeviceToDevice, c->stream));
    k_sadd_dev<<<vgrid, 256, 0, c->stream>>>(n, d + o, s_beta, -1.0, h + o);
    NSG_LAUNCH_CHECK(c);
  }
}

}  // namespace nsg
