// tri_layout_check.cpp — CPU test helper (tests/test_tri_layout.py): walks the level-order layout of the one-CTA ILU(0)
// triangular solves (navier-stokes-dealii_b200/csrc/nsg_tri_layout.h, the product's own header) the way k_ilu_solve_cta does and
// checks every invariant the kernel relies on:
//   * every factor entry sits in exactly one slot, padding slots have no source and read a position inside the window;
//   * a column flagged "in the window" is at most `window` positions behind the end of the reading level and its window
//     slot still holds that position when it is read; a column flagged "outside" has been written out AND that write-out
//     has completed (the kernel only knows a write-out is complete when it issues the next one);
//   * staged levels fit the rings next to their `depth` predecessors;
//   * the 8 lanes of a row, adding their partial sums in the order the kernel uses, give bitwise the sum a warp of 32 lanes
//     gives with stride-32 accumulation and the xor-shuffle tree (the order of the level-scheduled kernels).
// Test infrastructure: nothing here is linked into the product.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../navier-stokes-dealii_b200/csrc/nsg_tri_layout.h"

using namespace nsg;

namespace {

// sum of a row's products the way k_ilu_{l,u}solve_level / k_ilu_solve_sf do it: lane j accumulates entries j, j + 32, ...
// (fused multiply-add onto 0), then the xor-shuffle tree; lane 0's value
double warp_order_sum(const std::vector<double> &f, const std::vector<double> &y) {
  double lane[32];
  for (int j = 0; j < 32; ++j) lane[j] = 0.0;
  for (size_t p = 0; p < f.size(); ++p) lane[p % 32] = std::fma(f[p], y[p], lane[p % 32]);
  for (int o = 16; o > 0; o >>= 1) {
    double nxt[32];
    for (int j = 0; j < 32; ++j) nxt[j] = lane[j] + lane[j ^ o];
    std::memcpy(lane, nxt, sizeof lane);
  }
  return lane[0];
}

struct Walk {
  int64_t far_reads = 0, near_reads = 0, padding = 0, flushes = 0;
  int error = 0;  // first violated invariant
};

// One triangular solve over the layout. rhs/scale by position; returns the unknowns by position in `y`.
Walk walk(bool upper, const TriLayout &L, const TriLimits &lim, int flush, const double *fval, const std::vector<double> &rhs,
          const std::vector<double> &scale, int64_t n, const int64_t *rowptr, const int32_t *col, const int64_t *diag,
          std::vector<double> &y) {
  Walk w;
  const int32_t W = lim.window;
  y.assign(n, NAN);
  std::vector<double> win(W, NAN), written_out(n, NAN);
  std::vector<int64_t> win_pos(W, -1);
  int64_t flushed = 0, complete = 0;  // positions handed to the write-out engine / known to have landed
  std::vector<uint8_t> used(L.fsrc.size(), 0);
  auto fail = [&](int code) {
    if (!w.error) w.error = code;
  };
  for (int32_t l = 0; l < L.n_levels; ++l) {
    const TriRec a = L.info[l], b = L.info[l + 1];
    const int32_t k0 = a.w, k1 = b.w, slabs = a.y;
    // the compute warps of level l run while the producer may still be waiting for the previous write-out: a far read can only
    // count on what was known to have landed BEFORE this level's write-out decision
    const int64_t landed = complete;
    {  // the producer's write-out decision at the top of the level (k_ilu_solve_cta)
      const int32_t final_below = k0 & ~1;
      if (final_below - flushed >= flush) {
        complete = flushed;  // issuing a write-out first waits for the previous one
        for (int64_t p = flushed; p < final_below; ++p) {
          if (win_pos[p & (W - 1)] != p) fail(10);  // the window must still hold what is written out
          written_out[p] = win[p & (W - 1)];
        }
        flushed = final_below;
        ++w.flushes;
      }
    }
    if (a.z) {  // staged: fits the rings next to its predecessors
      const TriRec f = L.info[std::max(l - lim.depth, 0)];
      if (b.x - f.x > lim.ring_slots || b.w - f.w > lim.ring_rows) fail(20);
    }
    if (b.x - a.x != ((k1 - k0 + 3) / 4) * slabs * 32) fail(21);
    std::vector<double> level_y(k1 - k0);
    for (int32_t r = 0; r < k1 - k0; ++r) {
      const int32_t k = k0 + r, gi = r >> 2, sub = r & 3;
      const int64_t i = L.rows[k];
      if (L.pos[i] != k) fail(30);
      const int64_t p0 = upper ? diag[i] + 1 : rowptr[i], m = upper ? rowptr[i + 1] - diag[i] - 1 : diag[i] - rowptr[i];
      if (m > 8 * (int64_t)slabs) fail(31);
      // the 8 lanes of the row: lane c holds 4 partial sums, entry j = c + 8 s accumulates on t[s mod 4]
      double t[8][4];
      for (auto &row : t)
        for (double &v : row) v = 0.0;
      std::vector<double> fs, ys;  // the same entries in CSR order, for the 32-lane order
      for (int32_t s = 0; s < slabs; ++s)
        for (int c = 0; c < 8; ++c) {
          const size_t q = (size_t)a.x + ((size_t)gi * slabs + s) * 32 + 8 * sub + c;
          const TriRec e = L.slots[q];
          const int64_t j = 8 * s + c;
          double f = 0.0;
          if (j < m) {
            if (L.fsrc[q] != p0 + j || used[q]) fail(32);
            used[q] = 1;
            f = fval[L.fsrc[q]];
          } else {
            if (L.fsrc[q] != -1) fail(33);
            ++w.padding;
          }
          const int32_t cp = e.z >= 0 ? e.z : ~e.z;
          if (j < m && cp != L.pos[col[p0 + j]]) fail(34);
          if (cp >= k0 && k0 > 0) fail(35);  // a column of an earlier level
          double yv;
          if (e.z >= 0) {
            if (k1 - cp > W) fail(40);
            if (k0 > 0 && win_pos[cp & (W - 1)] != cp) fail(41);
            yv = k0 > 0 ? win[cp & (W - 1)] : 0.0;
            ++w.near_reads;
          } else {
            if (cp >= landed) fail(42);  // must have landed in global memory
            yv = written_out[cp];
            ++w.far_reads;
          }
          if (j < m && !(yv == y[cp])) fail(43);
          t[c][s & 3] = s < 4 ? f * yv : std::fma(f, yv, t[c][s & 3]);
          if (j < m) fs.push_back(f), ys.push_back(yv);
        }
      double lane8[8];
      for (int c = 0; c < 8; ++c) lane8[c] = (t[c][0] + t[c][2]) + (t[c][1] + t[c][3]);  // xor 16, xor 8
      for (int o = 4; o > 0; o >>= 1) {
        double nxt[8];
        for (int c = 0; c < 8; ++c) nxt[c] = lane8[c] + lane8[c ^ o];
        std::memcpy(lane8, nxt, sizeof lane8);
      }
      const double ref = warp_order_sum(fs, ys);
      if (std::memcmp(&ref, &lane8[0], sizeof ref) != 0 && !(ref == 0.0 && lane8[0] == 0.0)) fail(50);
      level_y[r] = upper ? std::fma(rhs[k], scale[k], -lane8[0]) : rhs[k] - lane8[0];
    }
    for (int32_t r = 0; r < k1 - k0; ++r) {  // the level's stores become visible at the next barrier
      const int32_t k = k0 + r;
      y[k] = level_y[r];
      win[k & (W - 1)] = level_y[r];
      win_pos[k & (W - 1)] = k;
    }
  }
  for (size_t q = 0; q + 1 < L.fsrc.size(); ++q)
    if (L.fsrc[q] >= 0 && !used[q]) fail(60);
  return w;
}
}  // namespace

extern "C" {
// val: a CSR matrix with its diagonal; the strictly lower part is taken as the unit lower factor, the strictly upper part as
// the (scaled) upper factor and 1 / diagonal as the inverse pivots - the roles they play in Ifpack's ILU(0) apply.
// y = U^-1 D L^-1 x by row.  stats[8] = {levels L, levels U, slots L, slots U, far reads, levels read in place, padding, flushes}.
// Returns 0 or the code of the first violated invariant (+100 for the backward solve).
int tri_layout_check(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *x, double *y_out,
                     int32_t window, int32_t ring_slots, int32_t ring_rows, int32_t depth, int32_t flush, int64_t *stats) {
  std::vector<int64_t> diag(n, -1);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p)
      if (col[p] == i) diag[i] = p;
  for (int64_t i = 0; i < n; ++i)
    if (diag[i] < 0) return -1;
  // dependency levels exactly as build_block (nsg_precond.cuh) computes them
  std::vector<int32_t> levL(n, 0), levU(n, 0);
  int32_t nL = 0, nU = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t l = 0;
    for (int64_t p = rowptr[i]; p < diag[i]; ++p) l = std::max(l, levL[col[p]] + 1);
    levL[i] = l, nL = std::max(nL, l + 1);
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    int32_t l = 0;
    for (int64_t p = diag[i] + 1; p < rowptr[i + 1]; ++p) l = std::max(l, levU[col[p]] + 1);
    levU[i] = l, nU = std::max(nU, l + 1);
  }
  auto bucket = [&](const std::vector<int32_t> &lev, int32_t nl, std::vector<int32_t> &ptr, std::vector<int32_t> &rows) {
    ptr.assign(nl + 1, 0);
    for (int64_t i = 0; i < n; ++i) ptr[lev[i] + 1]++;
    for (int32_t l = 0; l < nl; ++l) ptr[l + 1] += ptr[l];
    rows.resize(n);
    std::vector<int32_t> at(ptr.begin(), ptr.end() - 1);
    for (int64_t i = 0; i < n; ++i) rows[at[lev[i]]++] = (int32_t)i;
  };
  std::vector<int32_t> ptrL, ptrU, rowsL, rowsU;
  bucket(levL, nL, ptrL, rowsL);
  bucket(levU, nU, ptrU, rowsU);
  // the kernel's precondition (TRI_MAX_LEVEL_ROWS = window / 4, TRI_FLUSH = window / 16 in nsg_precond.cuh): a wider level or a
  // coarser write-out could overrun the window before it is written out
  for (int32_t l = 0; l < nL; ++l)
    if (ptrL[l + 1] - ptrL[l] > window / 4) return -3;
  for (int32_t l = 0; l < nU; ++l)
    if (ptrU[l + 1] - ptrU[l] > window / 4) return -3;
  if (flush > window / 16) return -3;
  const TriLimits lim{window, ring_slots, ring_rows, depth};
  TriLayout LL, LU;
  if (tri_layout(false, n, rowptr, col, diag.data(), rowsL, ptrL, nullptr, lim, LL)) return -2;
  if (tri_layout(true, n, rowptr, col, diag.data(), rowsU, ptrU, &LL.pos, lim, LU)) return -2;
  std::vector<double> rhs(n), scale(n, 1.0), yl, yu;
  for (int64_t k = 0; k < n; ++k) rhs[k] = x[LL.ra_src[k]];
  const Walk wl = walk(false, LL, lim, flush, val, rhs, scale, n, rowptr, col, diag.data(), yl);
  if (wl.error) return wl.error;
  for (int64_t k = 0; k < n; ++k) rhs[k] = yl[LU.ra_src[k]], scale[k] = 1.0 / val[diag[LU.rows[k]]];
  const Walk wu = walk(true, LU, lim, flush, val, rhs, scale, n, rowptr, col, diag.data(), yu);
  if (wu.error) return 100 + wu.error;
  for (int64_t k = 0; k < n; ++k) y_out[LU.rows[k]] = yu[k];
  if (stats) {
    stats[0] = nL, stats[1] = nU, stats[2] = LL.nq, stats[3] = LU.nq, stats[4] = wl.far_reads + wu.far_reads;
    stats[5] = LL.in_place + LU.in_place, stats[6] = wl.padding + wu.padding, stats[7] = wl.flushes + wu.flushes;
  }
  return 0;
}
}
