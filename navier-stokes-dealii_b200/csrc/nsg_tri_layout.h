// nsg_tri_layout.h — host side of the one-CTA ILU(0) triangular solves (k_ilu_solve_cta, nsg_precond.cuh): the level-order
// layout of a triangular factor.  Plain C++ (no CUDA) so that the CPU tests can build it too (tests/helpers/tri_layout_check.cpp
// walks the layout the way the kernel does and checks every invariant the kernel relies on).
//
// Rows are renumbered level by level ("positions"; inside a level by falling row length) and packed in groups of 4 consecutive
// positions, 8 lanes per row: entry j of a row belongs to lane j mod 8, slab j / 8.  A group's slabs are 32 records each (one
// per lane of the warp), every group of a level has the level's slab count (short rows are padded with zero factors).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace nsg {

struct TriRec {  // same layout as CUDA's int4
  int32_t x, y, z, w;
};
struct TriLimits {
  int32_t window;      // positions whose unknowns the shared-memory window holds (a power of two)
  int32_t ring_slots;  // slot records the ring holds (a power of two)
  int32_t ring_rows;   // row records the ring holds (a power of two)
  int32_t depth;       // levels staged ahead: levels l - depth .. l share the rings while level l is staged
};
struct TriLayout {
  int32_t n_levels = 0, in_place = 0;
  int64_t nq = 0;                // slots (entries incl. padding); `slots` and `fsrc` carry one spare record
  std::vector<TriRec> info;      // [n_levels+1] {first slot, slabs per group, 1 = staged through the rings / 0 = read in place, first position}
  std::vector<TriRec> slots;     // {factor (2 words, filled on the device), position of the entry's unknown or ~position once it has
                                 //  left the window, 0}
  std::vector<int64_t> fsrc;     // where the slot's entry sits in the factor array; -1 = padding (factor 0)
  std::vector<int32_t> rows;     // [n] row of each position
  std::vector<int32_t> pos;      // [n] position of each row
  std::vector<int32_t> ra_src;   // [n] by position - forward: the row (gathers the right-hand side); backward: its position in the
                                 //     forward layout (gathers the forward result)
};

// upper = false: the strictly lower part of every row (columns before diag[i]); true: the strictly upper part.
// rows_by_level / ptr: the rows sorted by dependency level and the level boundaries (a row's columns lie in earlier levels).
// Returns 0, or 1 (a row with more than 2^19 entries), 2 (more than 2^30 slots).
inline int tri_layout(bool upper, int64_t n, const int64_t *rowptr, const int32_t *col, const int64_t *diag,
                      const std::vector<int32_t> &rows_by_level, const std::vector<int32_t> &ptr, const std::vector<int32_t> *pos_fwd,
                      const TriLimits &lim, TriLayout &out) {
  const int32_t nl = (int32_t)ptr.size() - 1;
  auto len = [&](int64_t i) { return upper ? rowptr[i + 1] - diag[i] - 1 : diag[i] - rowptr[i]; };
  // positions: level by level, inside a level by falling row length
  std::vector<int32_t> &rows = out.rows, &pos = out.pos;
  rows = rows_by_level;
  for (int32_t l = 0; l < nl; ++l)
    std::stable_sort(rows.begin() + ptr[l], rows.begin() + ptr[l + 1], [&](int32_t x, int32_t y) { return len(x) > len(y); });
  pos.resize(n);
  for (int64_t k = 0; k < n; ++k) pos[rows[k]] = (int32_t)k;
  std::vector<TriRec> &info = out.info, &slots = out.slots;
  std::vector<int64_t> &fsrc = out.fsrc;
  info.assign(nl + 1, TriRec{0, 0, 0, 0});
  slots.clear(), fsrc.clear();
  out.ra_src.resize(n);
  for (int32_t l = 0; l < nl; ++l) {
    const int32_t k0 = ptr[l], k1 = ptr[l + 1];
    const int64_t longest = len(rows[k0]);  // the longest row of the level comes first
    if (longest > ((int64_t)1 << 19)) return 1;
    const int32_t slabs = (int32_t)((longest + 7) / 8);
    const size_t q0 = slots.size(), groups = (size_t)(k1 - k0 + 3) / 4;
    info[l] = TriRec{(int32_t)q0, slabs, 0, k0};
    const int32_t pad = std::max(k0 - 1, 0);  // padding (factor 0) reads the last unknown of the previous level
    slots.resize(q0 + groups * slabs * 32, TriRec{0, 0, k1 - pad <= lim.window ? pad : ~pad, 0});
    fsrc.resize(q0 + groups * slabs * 32, -1);
    if (slots.size() >= (size_t)1 << 30) return 2;
    for (int32_t r = 0; r < k1 - k0; ++r) {
      const int64_t i = rows[k0 + r];
      const int64_t p0 = upper ? diag[i] + 1 : rowptr[i], m = len(i);
      for (int64_t j = 0; j < m; ++j) {  // entry j: lane j mod 8 of the row, slab j / 8 of its group
        const int32_t cp = pos[col[p0 + j]];  // a row of an earlier level: cp < k0
        const size_t q = q0 + ((size_t)(r >> 2) * slabs + (size_t)(j >> 3)) * 32 + 8 * (r & 3) + (j & 7);
        slots[q].z = k1 - cp <= lim.window ? cp : ~cp;
        fsrc[q] = p0 + j;
      }
      out.ra_src[k0 + r] = upper ? (*pos_fwd)[i] : (int32_t)i;
    }
  }
  info[nl] = TriRec{(int32_t)slots.size(), 0, 0, (int32_t)n};
  out.in_place = 0;
  for (int32_t l = 0; l < nl; ++l) {  // levels l - depth .. l share the rings while level l is staged
    const int32_t f = std::max(l - lim.depth, 0);
    info[l].z = (info[l + 1].x - info[f].x <= lim.ring_slots && info[l + 1].w - info[f].w <= lim.ring_rows) ? 1 : 0;
    out.in_place += info[l].z ? 0 : 1;
  }
  out.n_levels = nl, out.nq = (int64_t)slots.size();
  slots.push_back(TriRec{0, 0, 0, 0}), fsrc.push_back(-1);
  return 0;
}

}  // namespace nsg
