// nst.cpp — host "topology" library (libnst.so), the deal.II-free restatement of what
// NavierStokesSolver::setup() hands to the per-Newton-step hot path.
//
// Reference sites restated here (paths relative to /root/reference):
//   GridIn::read_msh                          src/NavierStokesSolver.cpp:12-16
//   GridTools::partition_triangulation        src/NavierStokesSolver.cpp:18     (METIS -> RCB)
//   fullydistributed::create_triangulation    src/NavierStokesSolver.cpp:19-21  (owned cells + ghost layer)
//   FESystem(FE_SimplexP(2)^2, FE_SimplexP(1)) src/NavierStokesSolver.cpp:33-38  (15 DoFs per cell)
//   distribute_dofs + component_wise          src/NavierStokesSolver.cpp:64-73
//   make_sparsity_pattern x3                  src/NavierStokesSolver.cpp:107-158
//   interpolate_boundary_values               src/NavierStokesSolver.cpp:357-373
//
// Everything is plain host C++ (OpenMP where it pays); the arrays produced here are the
// one-time uploads of include/nsg.h.  This is product code, not the oracle.
#include "../../include/nst.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}

}  // namespace

struct nst_mesh {
  int64_t V = 0, T = 0, E = 0;
  std::vector<double> xy;           // [2V]
  std::vector<int32_t> cells;       // [3T]
  std::vector<int32_t> cell_edges;  // [3T]
  std::vector<int32_t> edge_v;      // [2E] orientation of the first cell that saw the edge
  std::vector<int32_t> edge_tag;    // [E]  boundary id, -1 interior
  std::vector<uint8_t> edge_nc;     // [E]  number of incident cells (1 = boundary)
  int64_t n_inverted = 0, n_boundary = 0;
};

// vector whose resize() leaves new elements uninitialised: the column arrays of the patterns (hundreds of millions of
// entries) are filled by the OpenMP loop right after; value-initialising them first was a serial 0.7 s per 160 M entries
template <class T>
struct DefaultInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  template <class U>
  void construct(U *p) noexcept {
    ::new (static_cast<void *>(p)) U;
  }
  template <class U, class... A>
  void construct(U *p, A &&...a) {
    ::new (static_cast<void *>(p)) U(std::forward<A>(a)...);
  }
};
using ColVec = std::vector<int32_t, DefaultInitAlloc<int32_t>>;

struct nst_dofs {
  int64_t n_u = 0, n_p = 0;
  int n_parts = 1;
  ColVec cell_dofs;                  // [15T] (1.1 GB at the bench size: first touched by the parallel loop that fills it)
  std::vector<int32_t> vertex_node;  // [V]
  std::vector<int32_t> edge_node;    // [E]
  std::vector<int32_t> vertex_p;     // [V]
  std::vector<int32_t> vertex_owner, edge_owner;
  std::vector<int64_t> part_n_u, part_n_p;  // owned counts
  std::vector<int64_t> u_off, p_off;        // prefix sums (n_parts+1)
};

struct nst_part {
  nst_part_info info{};
  std::vector<int64_t, DefaultInitAlloc<int64_t>> l2g;
  std::vector<int32_t> cell_ids;
  ColVec cell_dofs, cell_vertices;
  std::vector<double> xy;
  std::vector<uint8_t> cell_owned;
  std::vector<int64_t> jac_rowptr, pm_rowptr;
  ColVec jac_col, pm_col;
  std::vector<int32_t> neighbors;
  std::vector<int64_t> send_ptr, recv_ptr;
  std::vector<int32_t> send_idx, recv_idx;
  std::vector<int32_t> bface_cell, bface_face, bface_tag;
};

namespace {

// ---------------------------------------------------------------------------------------
// mesh construction from raw arrays
// ---------------------------------------------------------------------------------------
struct Line {
  int32_t a, b, tag;
};

// Finds the slot of (lo,hi) in the per-vertex sorted neighbour lists.
inline int64_t find_slot(const std::vector<int64_t> &off, const std::vector<int32_t> &his, int32_t lo,
                         int32_t hi) {
  const int32_t *b = his.data() + off[lo], *e = his.data() + off[lo + 1];
  const int32_t *p = std::lower_bound(b, e, hi);
  if (p == e || *p != hi) return -1;
  return p - his.data();
}

// ---- "number by first appearance", in parallel ------------------------------------------------------------------
// deal.II numbers lines, vertices and DoFs in the order a cell loop first meets them.  A serial loop does that in one
// pass; with 57 M (cell, line) slots per mesh level it was most of the set-up time.  The same numbers without the loop:
// every entity records the SMALLEST slot position that refers to it (atomic min); a slot is a first appearance iff it is
// that position; the entity's number is the count of first appearances before its slot (one scan).
inline void atomic_min(int64_t *addr, int64_t v) {
  int64_t cur = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(addr, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
}
inline void atomic_min32(int32_t *addr, int32_t v) {
  int32_t cur = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (v < cur && !__atomic_compare_exchange_n(addr, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
}
struct Tracer {  // NST_TRACE=1: wall time of the stages of the big set-up functions on stderr
  const char *fn;
  double t;
  bool on;
  explicit Tracer(const char *f) : fn(f), t(0), on(std::getenv("NST_TRACE") != nullptr) {
#ifdef _OPENMP
    t = omp_get_wtime();
#endif
  }
  void mark(const char *what) {
#ifdef _OPENMP
    if (on) {
      const double now = omp_get_wtime();
      std::fprintf(stderr, "[nst trace] %s: %s %.3f s\n", fn, what, now - t);
      t = now;
    }
#endif
  }
};
inline int n_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// Calls visit(pos, rank) for every position pos in [p0, p1) with is_first(pos), in ascending order of pos inside a chunk,
// where rank = rank0 + number of first positions in [p0, pos).  Returns rank0 + their total number.
template <class IsFirst, class Visit>
int64_t rank_first_positions(int64_t p0, int64_t p1, int64_t rank0, IsFirst is_first, Visit visit) {
  const int nt = n_threads();
  const int64_t n = p1 - p0, chunk = (n + nt - 1) / std::max(nt, 1);
  std::vector<int64_t> cnt(nt + 1, 0);
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nt; ++t) {
    int64_t k = 0;
    for (int64_t p = p0 + t * chunk, e = std::min(p1, p + chunk); p < e; ++p) k += is_first(p) ? 1 : 0;
    cnt[t + 1] = k;
  }
  for (int t = 0; t < nt; ++t) cnt[t + 1] += cnt[t];
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nt; ++t) {
    int64_t r = rank0 + cnt[t];
    for (int64_t p = p0 + t * chunk, e = std::min(p1, p + chunk); p < e; ++p)
      if (is_first(p)) visit(p, r++);
  }
  return rank0 + cnt[nt];
}

// Numbers the E unique edges ("slots") of m->cells by first appearance in (cell, local line) order and fills cell_edges,
// edge_v (orientation of the first cell that sees the edge) and edge_nc.  slot_of[3c + l] = slot of line l of cell c.
int number_edges(nst_mesh *m, const std::vector<int32_t> &slot_of, std::vector<int32_t> &slot_id) {
  const int64_t T = m->T, E = m->E;
  std::vector<int32_t> n_cells_of(E, 0);
  std::vector<int64_t> first_pos(E, INT64_MAX);
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < 3 * T; ++p) {
    const int32_t s = slot_of[p];
    atomic_min(&first_pos[s], p);
#pragma omp atomic
    n_cells_of[s]++;
  }
  int64_t too_many = 0;
#pragma omp parallel for reduction(+ : too_many) schedule(static)
  for (int64_t s = 0; s < E; ++s) too_many += n_cells_of[s] > 2 ? 1 : 0;
  if (too_many) {  // report the edge the cell loop would have stumbled over first
    std::vector<uint8_t> seen(E, 0);
    for (int64_t p = 0; p < 3 * T; ++p)
      if (++seen[slot_of[p]] == 3) {
        char buf[160];
        snprintf(buf, sizeof buf, "edge (%d,%d) is shared by more than two cells (overlapping surface entities?)",
                 m->cells[p], m->cells[3 * (p / 3) + (p % 3 + 1) % 3]);
        return fail(NST_ERR_TOPOLOGY, buf);
      }
  }
  slot_id.assign(E, -1);
  m->cell_edges.resize(3 * T);
  m->edge_v.resize(2 * E);
  m->edge_nc.assign(E, 0);
  rank_first_positions(
      0, 3 * T, 0, [&](int64_t p) { return first_pos[slot_of[p]] == p; },
      [&](int64_t p, int64_t id) {
        const int64_t c = p / 3;
        const int l = (int)(p % 3);
        slot_id[slot_of[p]] = (int32_t)id;
        m->edge_v[2 * id] = m->cells[3 * c + l];  // orientation of the first cell that sees the edge
        m->edge_v[2 * id + 1] = m->cells[3 * c + (l + 1) % 3];
        m->edge_nc[id] = (uint8_t)n_cells_of[slot_of[p]];
      });
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < 3 * T; ++p) m->cell_edges[p] = slot_id[slot_of[p]];
  return NST_OK;
}

int build_mesh(int64_t V, const double *xy, int64_t T, const int32_t *cells, int64_t n_lines,
               const Line *lines, nst_mesh *m) {
  Tracer tr("build_mesh");
  auto mark = [&](const char *w) { tr.mark(w); };
  m->V = V;
  m->T = T;
  m->xy.assign(xy, xy + 2 * V);
  m->cells.assign(cells, cells + 3 * T);
  int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
  for (int64_t i = 0; i < 3 * T; ++i) bad += (cells[i] < 0 || cells[i] >= V) ? 1 : 0;
  if (bad) return fail(NST_ERR_ARG, "cell vertex index out of range");
  // Negative-measure cells are inverted by swapping vertices 1 and 2.
  int64_t ninv = 0;
#pragma omp parallel for reduction(+ : ninv) schedule(static)
  for (int64_t c = 0; c < T; ++c) {
    int32_t *v = &m->cells[3 * c];
    const double *p0 = &m->xy[2 * v[0]], *p1 = &m->xy[2 * v[1]], *p2 = &m->xy[2 * v[2]];
    const double det = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]);
    if (det < 0) {
      std::swap(v[1], v[2]);
      ++ninv;
    }
  }
  m->n_inverted = ninv;
  mark("copy+orient");

  // unique edges: bucket half-edges by their lower vertex, dedupe per vertex, then number
  // by first appearance in (cell, local line) order: line 0=(v0,v1) 1=(v1,v2) 2=(v2,v0).
  std::vector<int64_t> off(V + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < T; ++c) {
    const int32_t *v = &m->cells[3 * c];
    for (int l = 0; l < 3; ++l) {
#pragma omp atomic
      off[std::min(v[l], v[(l + 1) % 3]) + 1]++;
    }
  }
  for (int64_t i = 0; i < V; ++i) off[i + 1] += off[i];
  mark("histogram");
  std::vector<int32_t> his(off[V]);
  {
    std::vector<int64_t> pos(off.begin(), off.end() - 1);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < T; ++c) {
      const int32_t *v = &m->cells[3 * c];
      for (int l = 0; l < 3; ++l) {
        const int32_t a = v[l], b = v[(l + 1) % 3];
        int64_t at;
#pragma omp atomic capture
        at = pos[std::min(a, b)]++;
        his[at] = std::max(a, b);  // the order inside a bucket does not matter: sorted below
      }
    }
  }
  mark("fill");
  // sort + unique per vertex (compacting in place)
  std::vector<int64_t> uoff(V + 1, 0);
#pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < V; ++i) {
    int32_t *b = his.data() + off[i], *e = his.data() + off[i + 1];
    std::sort(b, e);
    uoff[i + 1] = std::unique(b, e) - b;
  }
  for (int64_t i = 0; i < V; ++i) uoff[i + 1] += uoff[i];
  std::vector<int32_t> uhis(uoff[V]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < V; ++i)
    std::copy(his.data() + off[i], his.data() + off[i] + (uoff[i + 1] - uoff[i]), uhis.data() + uoff[i]);
  his.clear();
  his.shrink_to_fit();
  mark("sort+unique+copy");
  const int64_t E = uoff[V];
  if (E > (int64_t)INT32_MAX) return fail(NST_ERR_ARG, "more than 2^31-1 edges");
  m->E = E;
  // slot of every (cell, line) position
  std::vector<int32_t> slot_of(3 * (size_t)T);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < T; ++c) {
    const int32_t *v = &m->cells[3 * c];
    for (int l = 0; l < 3; ++l) {
      const int32_t a = v[l], b = v[(l + 1) % 3];
      slot_of[3 * c + l] = (int32_t)find_slot(uoff, uhis, std::min(a, b), std::max(a, b));
    }
  }
  mark("slots");
  std::vector<int32_t> slot_id;
  const int rc = number_edges(m, slot_of, slot_id);
  if (rc != NST_OK) return rc;
  mark("numbering");
  // boundary ids: boundary edges default to 0 (deal.II's default boundary_id), interior -1
  m->edge_tag.assign(E, -1);
  int64_t nb = 0;
#pragma omp parallel for reduction(+ : nb) schedule(static)
  for (int64_t e = 0; e < E; ++e)
    if (m->edge_nc[e] == 1) {
      m->edge_tag[e] = 0;
      ++nb;
    }
  m->n_boundary = nb;
  for (int64_t i = 0; i < n_lines; ++i) {
    const Line &L = lines[i];
    if (L.a < 0 || L.b < 0 || L.a >= V || L.b >= V) continue;
    const int64_t s = find_slot(uoff, uhis, std::min(L.a, L.b), std::max(L.a, L.b));
    if (s < 0) continue;  // a line that is not an edge of any kept triangle
    const int32_t e = slot_id[s];
    if (m->edge_nc[e] == 1) m->edge_tag[e] = L.tag;
  }
  mark("tags");
  return NST_OK;
}

// One level of red refinement WITHOUT rediscovering the edges: a child line is either half of a parent edge (slot 2e + which
// end) or one of the 3 interior lines of the parent cell (slot 2E + 3c + j), so its slot is a formula; the numbering by
// first appearance is then the same as build_mesh's.  Returns nullptr (and leaves *rc untouched) if a child cell came out
// with negative measure - build_mesh would swap its vertices, so the caller takes the general path.
nst_mesh *refine_direct(const nst_mesh *cur, int snap_id, double cx, double cy, double r, int *rc) {
  const int64_t V = cur->V, T = cur->T, E = cur->E;
  if (2 * E + 3 * T > (int64_t)INT32_MAX) return nullptr;
  auto *m = new nst_mesh;
  m->V = V + E, m->T = 4 * T, m->E = 2 * E + 3 * T;
  m->xy.resize(2 * (size_t)(V + E));
  std::copy(cur->xy.begin(), cur->xy.end(), m->xy.begin());
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e) {
    const int32_t a = cur->edge_v[2 * e], b = cur->edge_v[2 * e + 1];
    double x = 0.5 * (cur->xy[2 * a] + cur->xy[2 * b]), y = 0.5 * (cur->xy[2 * a + 1] + cur->xy[2 * b + 1]);
    if (snap_id >= 0 && cur->edge_tag[e] == snap_id) {
      const double dx = x - cx, dy = y - cy, d = std::sqrt(dx * dx + dy * dy);
      if (d > 0) {
        x = cx + dx * r / d;
        y = cy + dy * r / d;
      }
    }
    m->xy[2 * (V + e)] = x;
    m->xy[2 * (V + e) + 1] = y;
  }
  m->cells.resize(12 * (size_t)T);
  std::vector<int32_t> slot_of(12 * (size_t)T);
  int64_t inverted = 0;
#pragma omp parallel for reduction(+ : inverted) schedule(static)
  for (int64_t c = 0; c < T; ++c) {
    const int32_t *v = &cur->cells[3 * c];
    const int32_t *ce = &cur->cell_edges[3 * c];
    const int32_t m01 = (int32_t)(V + ce[0]), m12 = (int32_t)(V + ce[1]), m20 = (int32_t)(V + ce[2]);
    int32_t *o = &m->cells[12 * c];
    o[0] = v[0], o[1] = m01, o[2] = m20;
    o[3] = m01, o[4] = v[1], o[5] = m12;
    o[6] = m20, o[7] = m12, o[8] = v[2];
    o[9] = m01, o[10] = m12, o[11] = m20;
    for (int q = 0; q < 4; ++q) {
      const double *p0 = &m->xy[2 * o[3 * q]], *p1 = &m->xy[2 * o[3 * q + 1]], *p2 = &m->xy[2 * o[3 * q + 2]];
      inverted += (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p2[0] - p0[0]) * (p1[1] - p0[1]) < 0 ? 1 : 0;
    }
    // half of parent edge ce[k] at the end that is vertex w; interior line j of the parent
    auto half = [&](int k, int32_t w) { return 2 * ce[k] + (cur->edge_v[2 * ce[k]] == w ? 0 : 1); };
    const int32_t i0 = (int32_t)(2 * E + 3 * c), i1 = i0 + 1, i2 = i0 + 2;  // (m01,m20), (m01,m12), (m12,m20)
    int32_t *s = &slot_of[12 * c];
    s[0] = half(0, v[0]), s[1] = i0, s[2] = half(2, v[0]);   // child 0: (v0,m01) (m01,m20) (m20,v0)
    s[3] = half(0, v[1]), s[4] = half(1, v[1]), s[5] = i1;   // child 1: (m01,v1) (v1,m12) (m12,m01)
    s[6] = i2, s[7] = half(1, v[2]), s[8] = half(2, v[2]);   // child 2: (m20,m12) (m12,v2) (v2,m20)
    s[9] = i1, s[10] = i2, s[11] = i0;                       // child 3: (m01,m12) (m12,m20) (m20,m01)
  }
  if (inverted) {
    delete m;
    return nullptr;
  }
  m->n_inverted = 0;
  std::vector<int32_t> slot_id;
  *rc = number_edges(m, slot_of, slot_id);
  if (*rc != NST_OK) {
    delete m;
    return nullptr;
  }
  // the halves of a boundary edge inherit its id (the two "lines" the general path would have been given); the rest is interior
  m->edge_tag.assign(m->E, -1);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < E; ++e)
    if (cur->edge_nc[e] == 1) m->edge_tag[slot_id[2 * e]] = m->edge_tag[slot_id[2 * e + 1]] = cur->edge_tag[e];
  m->n_boundary = 2 * cur->n_boundary;
  return m;
}

// Drops vertices no triangle references, keeping file order (GridTools::delete_unused_vertices).
void compact_vertices(std::vector<double> &xy, std::vector<int32_t> &cells, std::vector<Line> &lines) {
  const int64_t V = (int64_t)xy.size() / 2;
  std::vector<int32_t> used(V, 0);
  for (int32_t v : cells) used[v] = 1;
  std::vector<int32_t> newid(V, -1);
  int32_t n = 0;
  for (int64_t i = 0; i < V; ++i)
    if (used[i]) newid[i] = n++;
  if (n == V) return;
  std::vector<double> nxy(2 * (size_t)n);
  for (int64_t i = 0; i < V; ++i)
    if (used[i]) {
      nxy[2 * newid[i]] = xy[2 * i];
      nxy[2 * newid[i] + 1] = xy[2 * i + 1];
    }
  xy.swap(nxy);
  for (auto &v : cells) v = newid[v];
  for (auto &L : lines) {
    L.a = newid[L.a];
    L.b = newid[L.b];
  }
}

bool seek_section(std::istream &in, const char *name) {
  std::string line;
  while (std::getline(in, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
    if (line == name) return true;
  }
  return false;
}

}  // namespace

extern "C" {

const char *nst_last_error(void) { return g_err.c_str(); }

int nst_mesh_create(int64_t n_vertices, const double *xy, int64_t n_cells, const int32_t *cells,
                    int64_t n_lines, const int32_t *line_v, const int32_t *line_tag, nst_mesh **out) {
  if (!xy || !cells || !out || n_vertices <= 0 || n_cells <= 0) return fail(NST_ERR_ARG, "null/empty mesh arrays");
  std::vector<Line> lines(n_lines);
  for (int64_t i = 0; i < n_lines; ++i) lines[i] = {line_v[2 * i], line_v[2 * i + 1], line_tag[i]};
  auto *m = new nst_mesh;
  const int rc = build_mesh(n_vertices, xy, n_cells, cells, n_lines, lines.data(), m);
  if (rc != NST_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return NST_OK;
}

int nst_mesh_read_msh(const char *path, int surface_entity, nst_mesh **out) {
  if (!path || !out) return fail(NST_ERR_ARG, "null argument");
  std::ifstream in(path);
  if (!in) return fail(NST_ERR_IO, std::string("cannot open ") + path);
  if (!seek_section(in, "$MeshFormat")) return fail(NST_ERR_FORMAT, "no $MeshFormat section");
  double version = 0;
  int ftype = 0, dsize = 0;
  in >> version >> ftype >> dsize;
  if (ftype != 0) return fail(NST_ERR_FORMAT, "binary .msh files are not supported");
  std::vector<double> xy;
  std::vector<int32_t> cells;
  std::vector<Line> lines;
  std::map<long, int32_t> node_index;  // gmsh node tag -> file-order index
  if (version >= 2.0 && version < 3.0) {
    if (!seek_section(in, "$Nodes")) return fail(NST_ERR_FORMAT, "no $Nodes section");
    long n = 0;
    in >> n;
    xy.resize(2 * (size_t)n);
    for (long i = 0; i < n; ++i) {
      long id;
      double x, y, z;
      in >> id >> x >> y >> z;
      node_index[id] = (int32_t)i;
      xy[2 * i] = x;
      xy[2 * i + 1] = y;
    }
    if (!seek_section(in, "$Elements")) return fail(NST_ERR_FORMAT, "no $Elements section");
    long ne = 0;
    in >> ne;
    for (long i = 0; i < ne; ++i) {
      long id;
      int type, ntags;
      in >> id >> type >> ntags;
      long phys = 0, elem = 0;
      for (int t = 0; t < ntags; ++t) {
        long v;
        in >> v;
        if (t == 0) phys = v;
        if (t == 1) elem = v;
      }
      static const int nn_of[] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
      if (type < 1 || type > 15) return fail(NST_ERR_FORMAT, "unsupported gmsh element type");
      const int nn = nn_of[type];
      long nd[27];
      for (int k = 0; k < nn; ++k) in >> nd[k];
      if (!in) return fail(NST_ERR_FORMAT, "truncated $Elements section");
      if (type == 1)
        lines.push_back({node_index.at(nd[0]), node_index.at(nd[1]), (int32_t)phys});
      else if (type == 2) {
        if (surface_entity >= 0 && elem != surface_entity) continue;
        for (int k = 0; k < 3; ++k) cells.push_back(node_index.at(nd[k]));
      }
    }
  } else if (version >= 4.0 && version < 5.0) {
    // $Entities: physical tag of every curve/surface (first tag; none -> 0)
    std::map<long, int32_t> curve_phys;
    if (!seek_section(in, "$Entities")) return fail(NST_ERR_FORMAT, "no $Entities section");
    long np, nc, ns, nv;
    in >> np >> nc >> ns >> nv;
    for (long i = 0; i < np; ++i) {
      long tag, nphys;
      double x, y, z;
      in >> tag >> x >> y >> z >> nphys;
      for (long k = 0; k < nphys; ++k) {
        long t;
        in >> t;
      }
    }
    for (long i = 0; i < nc; ++i) {
      long tag, nphys, nb;
      double b[6];
      in >> tag;
      for (double &d : b) in >> d;
      in >> nphys;
      int32_t first = 0;
      for (long k = 0; k < nphys; ++k) {
        long t;
        in >> t;
        if (k == 0) first = (int32_t)t;
      }
      in >> nb;
      for (long k = 0; k < nb; ++k) {
        long t;
        in >> t;
      }
      curve_phys[tag] = first;
    }
    if (!in) return fail(NST_ERR_FORMAT, "malformed $Entities section");
    if (!seek_section(in, "$Nodes")) return fail(NST_ERR_FORMAT, "no $Nodes section");
    long nblocks, nnodes, mintag, maxtag;
    in >> nblocks >> nnodes >> mintag >> maxtag;
    xy.reserve(2 * (size_t)nnodes);
    for (long b = 0; b < nblocks; ++b) {
      long edim, etag, param, nin;
      in >> edim >> etag >> param >> nin;
      std::vector<long> tags(nin);
      for (long k = 0; k < nin; ++k) in >> tags[k];
      for (long k = 0; k < nin; ++k) {
        double x, y, z;
        in >> x >> y >> z;
        if (param) {
          double u;
          for (long q = 0; q < edim; ++q) in >> u;
        }
        node_index[tags[k]] = (int32_t)(xy.size() / 2);
        xy.push_back(x);
        xy.push_back(y);
      }
    }
    if (!in) return fail(NST_ERR_FORMAT, "malformed $Nodes section");
    if (!seek_section(in, "$Elements")) return fail(NST_ERR_FORMAT, "no $Elements section");
    long nel, emin, emax;
    in >> nblocks >> nel >> emin >> emax;
    for (long b = 0; b < nblocks; ++b) {
      long edim, etag, nin;
      int type;
      in >> edim >> etag >> type >> nin;
      static const int nn_of[] = {0, 2, 3, 4, 4, 8, 6, 5, 3, 6, 9, 10, 27, 18, 14, 1};
      if (type < 1 || type > 15) return fail(NST_ERR_FORMAT, "unsupported gmsh element type");
      const int nn = nn_of[type];
      for (long k = 0; k < nin; ++k) {
        long id, nd[27];
        in >> id;
        for (int q = 0; q < nn; ++q) in >> nd[q];
        if (!in) return fail(NST_ERR_FORMAT, "truncated $Elements section");
        if (type == 1) {
          auto it = curve_phys.find(etag);
          lines.push_back({node_index.at(nd[0]), node_index.at(nd[1]), it == curve_phys.end() ? 0 : it->second});
        } else if (type == 2) {
          if (surface_entity >= 0 && etag != surface_entity) continue;
          for (int q = 0; q < 3; ++q) cells.push_back(node_index.at(nd[q]));
        }
      }
    }
  } else {
    return fail(NST_ERR_FORMAT, "unsupported $MeshFormat version (need 2.x or 4.x ASCII)");
  }
  if (cells.empty()) return fail(NST_ERR_FORMAT, "mesh has no triangles (3-D mesh or wrong surface entity?)");
  // lines that touch vertices no kept triangle uses are dropped before compaction
  {
    std::vector<uint8_t> used(xy.size() / 2, 0);
    for (int32_t v : cells) used[v] = 1;
    std::vector<Line> keep;
    for (const Line &L : lines)
      if (used[L.a] && used[L.b]) keep.push_back(L);
    lines.swap(keep);
  }
  compact_vertices(xy, cells, lines);
  auto *m = new nst_mesh;
  const int rc = build_mesh((int64_t)xy.size() / 2, xy.data(), (int64_t)cells.size() / 3, cells.data(),
                            (int64_t)lines.size(), lines.data(), m);
  if (rc != NST_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return NST_OK;
}

int nst_mesh_refine(const nst_mesh *in, int levels, int snap_id, double cx, double cy, double r,
                    nst_mesh **out) {
  if (!in || !out || levels < 0) return fail(NST_ERR_ARG, "bad argument");
  const nst_mesh *cur = in;
  nst_mesh *owned = nullptr;
  if (levels == 0) {
    *out = new nst_mesh(*in);
    return NST_OK;
  }
  for (int lev = 0; lev < levels; ++lev) {
    const int64_t V = cur->V, T = cur->T, E = cur->E;
    if (4 * T > (int64_t)700000000) {
      delete owned;
      return fail(NST_ERR_ARG, "refined mesh would exceed the 32-bit index range");
    }
    {
      int drc = NST_OK;
      nst_mesh *direct = refine_direct(cur, snap_id, cx, cy, r, &drc);
      if (drc != NST_OK) {
        delete owned;
        return drc;
      }
      if (direct) {
        delete owned;
        owned = direct;
        cur = direct;
        continue;
      }
    }
    std::vector<double> xy(2 * (size_t)(V + E));
    std::copy(cur->xy.begin(), cur->xy.end(), xy.begin());
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < E; ++e) {
      const int32_t a = cur->edge_v[2 * e], b = cur->edge_v[2 * e + 1];
      double x = 0.5 * (cur->xy[2 * a] + cur->xy[2 * b]), y = 0.5 * (cur->xy[2 * a + 1] + cur->xy[2 * b + 1]);
      if (snap_id >= 0 && cur->edge_tag[e] == snap_id) {
        const double dx = x - cx, dy = y - cy, d = std::sqrt(dx * dx + dy * dy);
        if (d > 0) {
          x = cx + dx * r / d;
          y = cy + dy * r / d;
        }
      }
      xy[2 * (V + e)] = x;
      xy[2 * (V + e) + 1] = y;
    }
    std::vector<int32_t> cells(12 * (size_t)T);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < T; ++c) {
      const int32_t *v = &cur->cells[3 * c];
      const int32_t *ce = &cur->cell_edges[3 * c];
      const int32_t m01 = (int32_t)(V + ce[0]), m12 = (int32_t)(V + ce[1]), m20 = (int32_t)(V + ce[2]);
      int32_t *o = &cells[12 * c];
      o[0] = v[0], o[1] = m01, o[2] = m20;
      o[3] = m01, o[4] = v[1], o[5] = m12;
      o[6] = m20, o[7] = m12, o[8] = v[2];
      o[9] = m01, o[10] = m12, o[11] = m20;
    }
    std::vector<Line> lines;
    lines.reserve(2 * (size_t)cur->n_boundary);
    for (int64_t e = 0; e < E; ++e)
      if (cur->edge_nc[e] == 1) {
        const int32_t a = cur->edge_v[2 * e], b = cur->edge_v[2 * e + 1];
        lines.push_back({a, (int32_t)(V + e), cur->edge_tag[e]});
        lines.push_back({(int32_t)(V + e), b, cur->edge_tag[e]});
      }
    auto *nm = new nst_mesh;
    const int rc = build_mesh(V + E, xy.data(), 4 * T, cells.data(), (int64_t)lines.size(), lines.data(), nm);
    delete owned;
    if (rc != NST_OK) {
      delete nm;
      return rc;
    }
    owned = nm;
    cur = nm;
  }
  *out = owned;
  return NST_OK;
}

int nst_mesh_tag_boundary_box(nst_mesh *m, int id_left, int id_right, int id_wall, int id_other) {
  if (!m) return fail(NST_ERR_ARG, "null mesh");
  double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
  for (int64_t v = 0; v < m->V; ++v) {
    xmin = std::min(xmin, m->xy[2 * v]);
    xmax = std::max(xmax, m->xy[2 * v]);
    ymin = std::min(ymin, m->xy[2 * v + 1]);
    ymax = std::max(ymax, m->xy[2 * v + 1]);
  }
  const double tol = 1e-9 * std::max(xmax - xmin, ymax - ymin);
  for (int64_t e = 0; e < m->E; ++e) {
    if (m->edge_nc[e] != 1) continue;
    const int32_t a = m->edge_v[2 * e], b = m->edge_v[2 * e + 1];
    const double xa = m->xy[2 * a], xb = m->xy[2 * b], ya = m->xy[2 * a + 1], yb = m->xy[2 * b + 1];
    if (std::fabs(xa - xmin) < tol && std::fabs(xb - xmin) < tol)
      m->edge_tag[e] = id_left;
    else if (std::fabs(xa - xmax) < tol && std::fabs(xb - xmax) < tol)
      m->edge_tag[e] = id_right;
    else if ((std::fabs(ya - ymin) < tol && std::fabs(yb - ymin) < tol) ||
             (std::fabs(ya - ymax) < tol && std::fabs(yb - ymax) < tol))
      m->edge_tag[e] = id_wall;
    else
      m->edge_tag[e] = id_other;
  }
  return NST_OK;
}

void nst_mesh_free(nst_mesh *m) { delete m; }
int64_t nst_mesh_n_vertices(const nst_mesh *m) { return m->V; }
int64_t nst_mesh_n_cells(const nst_mesh *m) { return m->T; }
int64_t nst_mesh_n_edges(const nst_mesh *m) { return m->E; }
int64_t nst_mesh_n_boundary_edges(const nst_mesh *m) { return m->n_boundary; }
int64_t nst_mesh_n_inverted(const nst_mesh *m) { return m->n_inverted; }
const double *nst_mesh_xy(const nst_mesh *m) { return m->xy.data(); }
const int32_t *nst_mesh_cells(const nst_mesh *m) { return m->cells.data(); }
const int32_t *nst_mesh_cell_edges(const nst_mesh *m) { return m->cell_edges.data(); }
const int32_t *nst_mesh_edge_vertices(const nst_mesh *m) { return m->edge_v.data(); }
const int32_t *nst_mesh_edge_tag(const nst_mesh *m) { return m->edge_tag.data(); }

int nst_mesh_boundary_faces(const nst_mesh *m, int32_t *out_cell, int32_t *out_face, int32_t *out_tag) {
  if (!m || !out_cell || !out_face || !out_tag) return fail(NST_ERR_ARG, "null argument");
  int64_t k = 0;
  for (int64_t c = 0; c < m->T; ++c)
    for (int f = 0; f < 3; ++f) {
      const int32_t e = m->cell_edges[3 * c + f];
      if (m->edge_nc[e] == 1) {
        out_cell[k] = (int32_t)c;
        out_face[k] = f;
        out_tag[k] = m->edge_tag[e];
        ++k;
      }
    }
  return NST_OK;
}

// ---------------------------------------------------------------------------------------
// partition: recursive coordinate bisection of cell centroids
// ---------------------------------------------------------------------------------------
// Recursive coordinate bisection of the cell centroids.  The part of a cell depends only on the SETS the bisections produce
// (the `mid - lo` smallest cells under the total order (coordinate, cell id)), so the two halves can run as OpenMP tasks and
// the result does not depend on the thread count.
static void rcb(const double *cen, std::vector<int32_t> &ids, int64_t lo, int64_t hi, int p0, int np, int32_t *part) {
  if (np == 1) {
    for (int64_t i = lo; i < hi; ++i) part[ids[i]] = p0;
    return;
  }
  double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
  for (int64_t i = lo; i < hi; ++i)
    for (int d = 0; d < 2; ++d) {
      const double x = cen[2 * (int64_t)ids[i] + d];
      mn[d] = std::min(mn[d], x);
      mx[d] = std::max(mx[d], x);
    }
  const int d = (mx[0] - mn[0] >= mx[1] - mn[1]) ? 0 : 1;
  const int npl = np / 2;
  const int64_t mid = lo + (hi - lo) * npl / np;
  std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
    const double xa = cen[2 * (int64_t)a + d], xb = cen[2 * (int64_t)b + d];
    return xa < xb || (xa == xb && a < b);
  });
#pragma omp task shared(ids) if (hi - lo > 100000)
  rcb(cen, ids, lo, mid, p0, npl, part);
#pragma omp task shared(ids) if (hi - lo > 100000)
  rcb(cen, ids, mid, hi, p0 + npl, np - npl, part);
#pragma omp taskwait
}

int nst_partition_rcb(const nst_mesh *m, int n_parts, int32_t *cell_part) {
  if (!m || !cell_part || n_parts < 1) return fail(NST_ERR_ARG, "bad argument");
  std::vector<int32_t> ids(m->T);
  std::iota(ids.begin(), ids.end(), 0);
  std::vector<double> cen(2 * (size_t)m->T);  // centroids once: the comparator of nth_element reads them many times
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < m->T; ++c) {
    const int32_t *v = &m->cells[3 * c];
    for (int d = 0; d < 2; ++d) cen[2 * c + d] = (m->xy[2 * v[0] + d] + m->xy[2 * v[1] + d] + m->xy[2 * v[2] + d]) / 3.0;
  }
#pragma omp parallel
#pragma omp single
  rcb(cen.data(), ids, 0, m->T, 0, n_parts, cell_part);
  return NST_OK;
}

// ---------------------------------------------------------------------------------------
// DoF numbering: first-visit order over cells (3 vertices, then 3 lines), then
// component_wise(block {0,0,1}) -> u dofs of P2 node n are 2n,2n+1; pressure dofs follow.
// ---------------------------------------------------------------------------------------
int nst_dofs_distribute(const nst_mesh *m, int n_parts, const int32_t *cell_part, nst_dofs **out) {
  if (!m || !out || n_parts < 1) return fail(NST_ERR_ARG, "bad argument");
  if (n_parts > 1 && !cell_part) return fail(NST_ERR_ARG, "cell_part required for n_parts > 1");
  auto *d = new nst_dofs;
  Tracer tr("dofs");
  d->n_parts = n_parts;
  const int64_t V = m->V, E = m->E, T = m->T;
  d->vertex_node.assign(V, -1);
  d->vertex_p.assign(V, -1);
  d->edge_node.assign(E, -1);
  d->vertex_owner.assign(V, 0);
  d->edge_owner.assign(E, 0);
  std::vector<int64_t> order(T);  // cells grouped by part, file order inside a part
  std::vector<int64_t> part_begin;  // [n_parts+1] where the cells of a part start in `order`
  if (n_parts > 1) {
    std::fill(d->vertex_owner.begin(), d->vertex_owner.end(), INT32_MAX);
    std::fill(d->edge_owner.begin(), d->edge_owner.end(), INT32_MAX);
    std::vector<int64_t> cnt(n_parts + 1, 0);
    int64_t out_of_range = 0;
#pragma omp parallel for reduction(+ : out_of_range) schedule(static)
    for (int64_t c = 0; c < T; ++c) {
      const int32_t p = cell_part[c];
      if (p < 0 || p >= n_parts) {
        ++out_of_range;
        continue;
      }
      for (int k = 0; k < 3; ++k) {  // an entity belongs to the lowest part among its cells
        atomic_min32(&d->vertex_owner[m->cells[3 * c + k]], p);
        atomic_min32(&d->edge_owner[m->cell_edges[3 * c + k]], p);
      }
    }
    if (out_of_range) {
      delete d;
      return fail(NST_ERR_ARG, "cell_part entry out of range");
    }
    for (int64_t c = 0; c < T; ++c) cnt[cell_part[c] + 1]++;
    for (int p = 0; p < n_parts; ++p) cnt[p + 1] += cnt[p];
    part_begin.assign(cnt.begin(), cnt.end());
    for (int64_t c = 0; c < T; ++c) order[cnt[cell_part[c]]++] = c;
  } else {
    std::iota(order.begin(), order.end(), 0);
    part_begin = {0, T};
  }
  tr.mark("owners + order");
  d->part_n_u.assign(n_parts, 0);
  d->part_n_p.assign(n_parts, 0);
  // pass 1: per-part first-visit ranks (node rank and pressure rank local to the part).  The cell loop of a part visits, per
  // cell, its 3 vertices and then its 3 lines: position 6 i + k of the i-th cell in `order`; an entity is numbered at the first
  // position that refers to it among the cells of its owner part (see rank_first_positions).
  std::vector<int64_t> nn(n_parts, 0), np(n_parts, 0);
  {
    std::vector<int64_t> first_v(V, INT64_MAX), first_e(E, INT64_MAX);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < T; ++i) {
      const int64_t c = order[i];
      const int32_t p = n_parts > 1 ? cell_part[c] : 0;
      for (int k = 0; k < 3; ++k) {
        const int32_t v = m->cells[3 * c + k], e = m->cell_edges[3 * c + k];
        if (d->vertex_owner[v] == p) atomic_min(&first_v[v], 6 * i + k);
        if (d->edge_owner[e] == p) atomic_min(&first_e[e], 6 * i + 3 + k);
      }
    }
    auto is_first = [&](int64_t pos) {
      const int64_t c = order[pos / 6];
      const int k = (int)(pos % 6);
      return k < 3 ? first_v[m->cells[3 * c + k]] == pos : first_e[m->cell_edges[3 * c + k - 3]] == pos;
    };
    const int nt = n_threads();
    for (int p = 0; p < n_parts; ++p) {  // both ranks (node, pressure vertex) in one counting and one numbering pass
      const int64_t b = 6 * part_begin[p], e = 6 * part_begin[p + 1], chunk = (e - b + nt - 1) / std::max(nt, 1);
      std::vector<int64_t> cn(nt + 1, 0), cp(nt + 1, 0);
#pragma omp parallel for schedule(static, 1)
      for (int t = 0; t < nt; ++t) {
        int64_t kn = 0, kp = 0;
        for (int64_t pos = std::min(e, b + t * chunk), pe = std::min(e, pos + chunk); pos < pe; ++pos)
          if (is_first(pos)) ++kn, kp += pos % 6 < 3 ? 1 : 0;
        cn[t + 1] = kn, cp[t + 1] = kp;
      }
      for (int t = 0; t < nt; ++t) cn[t + 1] += cn[t], cp[t + 1] += cp[t];
#pragma omp parallel for schedule(static, 1)
      for (int t = 0; t < nt; ++t) {
        int64_t rn = cn[t], rp = cp[t];
        for (int64_t pos = std::min(e, b + t * chunk), pe = std::min(e, pos + chunk); pos < pe; ++pos)
          if (is_first(pos)) {
            const int64_t c = order[pos / 6];
            const int k = (int)(pos % 6);
            if (k < 3) {
              d->vertex_node[m->cells[3 * c + k]] = (int32_t)rn++;
              d->vertex_p[m->cells[3 * c + k]] = (int32_t)rp++;
            } else {
              d->edge_node[m->cell_edges[3 * c + k - 3]] = (int32_t)rn++;
            }
          }
      }
      nn[p] = cn[nt], np[p] = cp[nt];
    }
  }
  tr.mark("first-visit numbering");
  d->u_off.assign(n_parts + 1, 0);
  d->p_off.assign(n_parts + 1, 0);
  std::vector<int64_t> node_off(n_parts + 1, 0);
  for (int p = 0; p < n_parts; ++p) {
    d->part_n_u[p] = 2 * nn[p];
    d->part_n_p[p] = np[p];
    node_off[p + 1] = node_off[p] + nn[p];
    d->u_off[p + 1] = d->u_off[p] + 2 * nn[p];
    d->p_off[p + 1] = d->p_off[p] + np[p];
  }
  d->n_u = d->u_off[n_parts];
  d->n_p = d->p_off[n_parts];
  if (d->n_u + d->n_p > (int64_t)INT32_MAX) {
    delete d;
    return fail(NST_ERR_ARG, "more than 2^31-1 DoFs: 32-bit global_dof_index overflow");
  }
  // pass 2: shift part-local ranks to the block-wise, part-major global numbering
  if (n_parts > 1) {
    for (int64_t v = 0; v < V; ++v) {
      const int32_t p = d->vertex_owner[v];
      d->vertex_node[v] += (int32_t)node_off[p];
      d->vertex_p[v] += (int32_t)d->p_off[p];
    }
    for (int64_t e = 0; e < E; ++e) d->edge_node[e] += (int32_t)node_off[d->edge_owner[e]];
  }
  tr.mark("offsets");
  d->cell_dofs.resize(15 * (size_t)T);
  const int32_t nu = (int32_t)d->n_u;
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < T; ++c) {
    int32_t *o = &d->cell_dofs[15 * c];
    for (int k = 0; k < 3; ++k) {
      const int32_t v = m->cells[3 * c + k];
      o[3 * k] = 2 * d->vertex_node[v];
      o[3 * k + 1] = 2 * d->vertex_node[v] + 1;
      o[3 * k + 2] = nu + d->vertex_p[v];
      const int32_t e = m->cell_edges[3 * c + k];
      o[9 + 2 * k] = 2 * d->edge_node[e];
      o[9 + 2 * k + 1] = 2 * d->edge_node[e] + 1;
    }
  }
  tr.mark("cell_dofs");
  *out = d;
  return NST_OK;
}

void nst_dofs_free(nst_dofs *d) { delete d; }
int64_t nst_dofs_n_u(const nst_dofs *d) { return d->n_u; }
int64_t nst_dofs_n_p(const nst_dofs *d) { return d->n_p; }
const int32_t *nst_dofs_cell_dofs(const nst_dofs *d) { return d->cell_dofs.data(); }
const int32_t *nst_dofs_vertex_node(const nst_dofs *d) { return d->vertex_node.data(); }
const int32_t *nst_dofs_edge_node(const nst_dofs *d) { return d->edge_node.data(); }
const int32_t *nst_dofs_vertex_p(const nst_dofs *d) { return d->vertex_p.data(); }
const int64_t *nst_dofs_part_n_u(const nst_dofs *d) { return d->part_n_u.data(); }
const int64_t *nst_dofs_part_n_p(const nst_dofs *d) { return d->part_n_p.data(); }

}  // extern "C"

namespace {

// dof -> cells CSR restricted to rows < n_rows (the order of the cells of a dof is irrelevant to its users)
struct DofCells {
  std::vector<int64_t> ptr;
  std::vector<int32_t> cell;
};
static void dof_to_cells(int64_t n_rows, int64_t n_cells, const int32_t *cell_dofs, DofCells &dc) {
  std::vector<int64_t> &dptr = dc.ptr;
  dptr.assign(n_rows + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < 15 * n_cells; ++i)
    if (cell_dofs[i] < n_rows) {
#pragma omp atomic
      dptr[cell_dofs[i] + 1]++;
    }
  for (int64_t r = 0; r < n_rows; ++r) dptr[r + 1] += dptr[r];
  dc.cell.resize(dptr[n_rows]);
  std::vector<int64_t> pos(dptr.begin(), dptr.end() - 1);
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < n_cells; ++c)
    for (int k = 0; k < 15; ++k) {
      const int32_t g = cell_dofs[15 * c + k];
      if (g < n_rows) {
        int64_t at;
#pragma omp atomic capture
        at = pos[g]++;
        dc.cell[at] = (int32_t)c;
      }
    }
}

// Generic CSR pattern of the rows [0,n_rows) from 15-dof cell lists; `is_p(id)` tells the block
// of a (local) dof id. kind 0: all couplings; 1: all but p-p; 2: p-p only. Columns ascending.
template <class IsP>
void build_pattern(int64_t n_rows, int64_t n_cells, const int32_t *cell_dofs, int kind, IsP is_p,
                   std::vector<int64_t> &rowptr, ColVec *col, const DofCells *shared = nullptr) {
  DofCells own;
  if (!shared) {
    dof_to_cells(n_rows, n_cells, cell_dofs, own);
    shared = &own;
  }
  const std::vector<int64_t> &dptr = shared->ptr;
  const std::vector<int32_t> &dcell = shared->cell;
  // The two dofs of a velocity node are consecutive ids everywhere (global, owned-local and ghost-local numbering), so a
  // row is assembled from KEYS - the x-dof of every velocity node and every pressure dof of the patch (9 per cell
  // instead of 15 dofs) - which are sorted, made unique and then expanded (x-dof -> x-dof, y-dof).
  static const int UX[6] = {0, 3, 6, 9, 11, 13}, PD[3] = {2, 5, 8};
  auto row_cols = [&](int64_t r, int32_t *buf, int32_t *keys) -> int {
    const bool rp = is_p((int32_t)r);
    if (kind == 2 && !rp) return 0;  // the pressure-mass pattern has entries in pressure rows only
    int nk = 0;
    for (int64_t q = dptr[r]; q < dptr[r + 1]; ++q) {
      const int32_t *cd = cell_dofs + 15 * (int64_t)dcell[q];
      if (kind != 2)
        for (int k = 0; k < 6; ++k) keys[nk++] = cd[UX[k]];
      if (!(kind == 1 && rp) && !(kind == 2 && !rp))
        for (int k = 0; k < 3; ++k) keys[nk++] = cd[PD[k]];
    }
    std::sort(keys, keys + nk);
    nk = (int)(std::unique(keys, keys + nk) - keys);
    int n = 0;
    for (int j = 0; j < nk; ++j) {
      buf[n++] = keys[j];
      if (!is_p(keys[j])) buf[n++] = keys[j] + 1;
    }
    return n;
  };
  rowptr.assign(n_rows + 1, 0);
#pragma omp parallel
  {
    std::vector<int32_t> buf(15 * 64), keys(9 * 64);
#pragma omp for schedule(dynamic, 8192)
    for (int64_t r = 0; r < n_rows; ++r) {
      const size_t need = 15 * (size_t)(dptr[r + 1] - dptr[r]);
      if (buf.size() < need) buf.resize(need), keys.resize(need);
      rowptr[r + 1] = row_cols(r, buf.data(), keys.data());
    }
  }
  for (int64_t r = 0; r < n_rows; ++r) rowptr[r + 1] += rowptr[r];
  if (!col) return;
  col->resize(rowptr[n_rows]);
#pragma omp parallel
  {
    std::vector<int32_t> buf(15 * 64), keys(9 * 64);
#pragma omp for schedule(dynamic, 8192)
    for (int64_t r = 0; r < n_rows; ++r) {
      const size_t need = 15 * (size_t)(dptr[r + 1] - dptr[r]);
      if (buf.size() < need) buf.resize(need), keys.resize(need);
      const int n = row_cols(r, buf.data(), keys.data());
      std::copy(buf.data(), buf.data() + n, col->data() + rowptr[r]);
    }
  }
}

}  // namespace

extern "C" {

int nst_sparsity(const nst_mesh *m, const nst_dofs *d, int kind, int64_t *nnz, int64_t *rowptr, int32_t *col) {
  if (!m || !d || !nnz || kind < 0 || kind > 2) return fail(NST_ERR_ARG, "bad argument");
  const int64_t N = d->n_u + d->n_p;
  const int32_t nu = (int32_t)d->n_u;
  std::vector<int64_t> rp;
  ColVec cc;
  build_pattern(N, m->T, d->cell_dofs.data(), kind, [nu](int32_t g) { return g >= nu; }, rp,
                (rowptr && col) ? &cc : nullptr);
  *nnz = rp[N];
  if (rowptr && col) {
    std::copy(rp.begin(), rp.end(), rowptr);
    std::copy(cc.begin(), cc.end(), col);
  }
  return NST_OK;
}

int nst_dofs_support_points(const nst_mesh *m, const nst_dofs *d, double *xy) {
  if (!m || !d || !xy) return fail(NST_ERR_ARG, "null argument");
  const int64_t nu = d->n_u;
#pragma omp parallel for schedule(static)
  for (int64_t v = 0; v < m->V; ++v) {
    const double x = m->xy[2 * v], y = m->xy[2 * v + 1];
    const int64_t n = d->vertex_node[v];
    for (int a = 0; a < 2; ++a) {
      xy[2 * (2 * n + a)] = x;
      xy[2 * (2 * n + a) + 1] = y;
    }
    const int64_t p = nu + d->vertex_p[v];
    xy[2 * p] = x;
    xy[2 * p + 1] = y;
  }
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < m->E; ++e) {
    const int32_t a = m->edge_v[2 * e], b = m->edge_v[2 * e + 1];
    const double x = 0.5 * (m->xy[2 * a] + m->xy[2 * b]), y = 0.5 * (m->xy[2 * a + 1] + m->xy[2 * b + 1]);
    const int64_t n = d->edge_node[e];
    for (int c = 0; c < 2; ++c) {
      xy[2 * (2 * n + c)] = x;
      xy[2 * (2 * n + c) + 1] = y;
    }
  }
  return NST_OK;
}

// interpolate_boundary_values (cpp:357-373): successive calls accumulate into one ordered
// map; inside a call faces are visited in (cell, face) order and later visits overwrite.
int nst_dirichlet_values(const nst_mesh *m, const nst_dofs *d, int n_calls, const int32_t *call_ptr,
                         const int32_t *ids, const int32_t *is_inlet, const nst_inlet_params *inlet,
                         int64_t *n_out, int32_t *out_dof, double *out_val) {
  if (!m || !d || !call_ptr || !ids || !is_inlet || !inlet || !n_out) return fail(NST_ERR_ARG, "null argument");
  std::map<int32_t, double> bv;
  auto inlet_ux = [&](double y) {
    const double yy = y - inlet->y0;
    return 4. * inlet->u_m * yy * (inlet->H - yy) * inlet->time_factor / (inlet->H * inlet->H);
  };
  // the boundary faces in (cell, face) order: one parallel pass over the cells instead of one serial pass per call
  std::vector<int64_t> bfaces;  // 3 c + f
  {
    const int nt = n_threads();
    const int64_t ch = (m->T + nt - 1) / std::max(nt, 1);
    std::vector<std::vector<int64_t>> found(nt);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t)
      for (int64_t c = std::min(m->T, t * ch), ce = std::min(m->T, c + ch); c < ce; ++c)
        for (int f = 0; f < 3; ++f)
          if (m->edge_nc[m->cell_edges[3 * c + f]] == 1) found[t].push_back(3 * c + f);
    for (int t = 0; t < nt; ++t) bfaces.insert(bfaces.end(), found[t].begin(), found[t].end());
  }
  for (int call = 0; call < n_calls; ++call) {
    for (const int64_t cf : bfaces) {
        const int64_t c = cf / 3;
        const int f = (int)(cf % 3);
        const int32_t e = m->cell_edges[3 * c + f];
        int hit = -1;
        for (int32_t q = call_ptr[call]; q < call_ptr[call + 1]; ++q)
          if (ids[q] == m->edge_tag[e]) hit = q;
        if (hit < 0) continue;
        const int32_t va = m->cells[3 * c + f], vb = m->cells[3 * c + (f + 1) % 3];
        const double ya = m->xy[2 * va + 1], yb = m->xy[2 * vb + 1];
        const double ys[3] = {ya, yb, 0.5 * (ya + yb)};
        const int32_t nodes[3] = {d->vertex_node[va], d->vertex_node[vb], d->edge_node[e]};
        for (int k = 0; k < 3; ++k) {
          bv[2 * nodes[k]] = is_inlet[hit] ? inlet_ux(ys[k]) : 0.0;
          bv[2 * nodes[k] + 1] = 0.0;
        }
      }
  }
  if (out_dof && out_val) {
    if (*n_out < (int64_t)bv.size()) return fail(NST_ERR_ARG, "output buffers too small");
    int64_t k = 0;
    for (const auto &kv : bv) {
      out_dof[k] = kv.first;
      out_val[k] = kv.second;
      ++k;
    }
  }
  *n_out = (int64_t)bv.size();
  return NST_OK;
}

// ---------------------------------------------------------------------------------------
// one rank's local problem
// ---------------------------------------------------------------------------------------
int nst_part_build(const nst_mesh *m, const nst_dofs *d, int n_parts, const int32_t *cell_part, int rank,
                   nst_part **out) {
  return nst_part_build_ex(m, d, n_parts, cell_part, rank, 0, out);
}

int nst_part_build_ex(const nst_mesh *m, const nst_dofs *d, int n_parts, const int32_t *cell_part, int rank, int flags,
                      nst_part **out) {
  if (!m || !d || !out || n_parts != d->n_parts || rank < 0 || rank >= n_parts)
    return fail(NST_ERR_ARG, "bad argument (n_parts must match nst_dofs_distribute)");
  if (n_parts > 1 && !cell_part) return fail(NST_ERR_ARG, "cell_part required");
  auto *P = new nst_part;
  Tracer tr("part");
  const int64_t T = m->T, nu = d->n_u;
  const int64_t u0 = d->u_off[rank], u1 = d->u_off[rank + 1], p0 = d->p_off[rank], p1 = d->p_off[rank + 1];
  const int64_t n_own_u = u1 - u0, n_own_p = p1 - p0, n_own = n_own_u + n_own_p;
  auto owned_g = [&](int32_t g) { return g < nu ? (g >= u0 && g < u1) : (g - nu >= p0 && g - nu < p1); };
  auto owner_of = [&](int32_t g) {
    const std::vector<int64_t> &off = g < nu ? d->u_off : d->p_off;
    const int64_t x = g < nu ? g : g - nu;
    return (int)(std::upper_bound(off.begin(), off.end(), x) - off.begin()) - 1;
  };
  // Order-preserving parallel filter: calls emit(thread_local_sink, i) for i in [0, n) in chunks and concatenates the chunks in
  // order, so the result equals the serial loop's.
  const int nt = n_threads();
  auto chunk_of = [nt](int64_t n, int t, int64_t &b, int64_t &e) {
    const int64_t ch = (n + nt - 1) / std::max(nt, 1);
    b = std::min(n, t * ch), e = std::min(n, b + ch);
  };
  // local cells: every cell touching an owned DoF, in global order
  {
    std::vector<std::vector<int32_t>> found(nt);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
      int64_t b, e;
      chunk_of(T, t, b, e);
      for (int64_t c = b; c < e; ++c) {
        const int32_t *cd = &d->cell_dofs[15 * c];
        bool any = false;
        for (int k = 0; k < 15 && !any; ++k) any = owned_g(cd[k]);
        if (any) found[t].push_back((int32_t)c);
      }
    }
    for (int t = 0; t < nt; ++t) P->cell_ids.insert(P->cell_ids.end(), found[t].begin(), found[t].end());
  }
  const int64_t nc = (int64_t)P->cell_ids.size();
  tr.mark("cell ids");
  // ghosts
  std::vector<int32_t> gu, gp;
  auto uniq = [](std::vector<int32_t> &v) {
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
  };
  if (n_parts > 1) {
    std::vector<std::vector<int32_t>> tu(nt), tp(nt);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
      int64_t b, e;
      chunk_of(nc, t, b, e);
      for (int64_t i = b; i < e; ++i) {
        const int32_t *cd = &d->cell_dofs[15 * (int64_t)P->cell_ids[i]];
        for (int k = 0; k < 15; ++k)
          if (!owned_g(cd[k])) (cd[k] < nu ? tu[t] : tp[t]).push_back(cd[k]);
      }
      uniq(tu[t]), uniq(tp[t]);
    }
    for (int t = 0; t < nt; ++t) gu.insert(gu.end(), tu[t].begin(), tu[t].end()), gp.insert(gp.end(), tp[t].begin(), tp[t].end());
  }
  uniq(gu);
  uniq(gp);
  const int64_t n_gu = (int64_t)gu.size(), n_gp = (int64_t)gp.size();
  const int64_t n_loc = n_own + n_gu + n_gp;
  P->l2g.resize(n_loc);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_own_u; ++i) P->l2g[i] = u0 + i;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_own_p; ++i) P->l2g[n_own_u + i] = nu + p0 + i;
  for (int64_t i = 0; i < n_gu; ++i) P->l2g[n_own + i] = gu[i];
  for (int64_t i = 0; i < n_gp; ++i) P->l2g[n_own + n_gu + i] = gp[i];
  auto g2l = [&](int32_t g) -> int32_t {
    if (g < nu) {
      if (g >= u0 && g < u1) return (int32_t)(g - u0);
      return (int32_t)(n_own + (std::lower_bound(gu.begin(), gu.end(), g) - gu.begin()));
    }
    if (g - nu >= p0 && g - nu < p1) return (int32_t)(n_own_u + (g - nu - p0));
    return (int32_t)(n_own + n_gu + (std::lower_bound(gp.begin(), gp.end(), g) - gp.begin()));
  };
  tr.mark("ghosts + l2g");
  // local cells, vertices, dofs; local vertices are numbered by first appearance in the local cell loop
  P->cell_dofs.resize(15 * (size_t)nc);
  P->cell_vertices.resize(3 * (size_t)nc);
  P->cell_owned.resize(nc);
  std::vector<int32_t> vloc(m->V, -1);
  int32_t nv = 0;
  {
    std::vector<int64_t> first_v(m->V, INT64_MAX);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nc; ++i) {
      const int64_t c = P->cell_ids[i];
      for (int k = 0; k < 15; ++k) P->cell_dofs[15 * i + k] = g2l(d->cell_dofs[15 * c + k]);
      for (int k = 0; k < 3; ++k) atomic_min(&first_v[m->cells[3 * c + k]], 3 * i + k);
      P->cell_owned[i] = (n_parts == 1 || cell_part[c] == rank) ? 1 : 0;
    }
    auto vertex_at = [&](int64_t pos) { return m->cells[3 * (int64_t)P->cell_ids[pos / 3] + pos % 3]; };
    nv = (int32_t)rank_first_positions(
        0, 3 * nc, 0, [&](int64_t pos) { return first_v[vertex_at(pos)] == pos; },
        [&](int64_t pos, int64_t rank) { vloc[vertex_at(pos)] = (int32_t)rank; });
    P->xy.resize(2 * (size_t)nv);
#pragma omp parallel for schedule(static)
    for (int64_t pos = 0; pos < 3 * nc; ++pos) {
      const int32_t v = vertex_at(pos), l = vloc[v];
      P->cell_vertices[pos] = l;
      if (first_v[v] == pos) P->xy[2 * l] = m->xy[2 * v], P->xy[2 * l + 1] = m->xy[2 * v + 1];
    }
  }
  tr.mark("local cells + vertices");
  {  // boundary faces in local cell order
    std::vector<std::vector<int32_t>> fc(nt), ff(nt), ft(nt);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
      int64_t b, e;
      chunk_of(nc, t, b, e);
      for (int64_t i = b; i < e; ++i) {
        const int64_t c = P->cell_ids[i];
        for (int f = 0; f < 3; ++f) {
          const int32_t ed = m->cell_edges[3 * c + f];
          if (m->edge_nc[ed] == 1) fc[t].push_back((int32_t)i), ff[t].push_back(f), ft[t].push_back(m->edge_tag[ed]);
        }
      }
    }
    for (int t = 0; t < nt; ++t) {
      P->bface_cell.insert(P->bface_cell.end(), fc[t].begin(), fc[t].end());
      P->bface_face.insert(P->bface_face.end(), ff[t].begin(), ff[t].end());
      P->bface_tag.insert(P->bface_tag.end(), ft[t].begin(), ft[t].end());
    }
  }
  tr.mark("boundary faces");
  // local patterns of the owned rows (columns ascending in local ids: owned first, then ghosts)
  const int32_t a0 = (int32_t)n_own_u, a1 = (int32_t)n_own, a2 = (int32_t)(n_own + n_gu);
  auto is_p = [a0, a1, a2](int32_t l) { return (l >= a0 && l < a1) || l >= a2; };
  if (flags & NST_PART_NO_PATTERNS) {  // the device builds them from cell_dofs (nsg_set_pattern_from_cells): the row pointers stay empty
  } else {
    DofCells dc;  // shared by the two patterns
    dof_to_cells(n_own, nc, P->cell_dofs.data(), dc);
    build_pattern(n_own, nc, P->cell_dofs.data(), 0, is_p, P->jac_rowptr, &P->jac_col, &dc);
    build_pattern(n_own, nc, P->cell_dofs.data(), 2, is_p, P->pm_rowptr, &P->pm_col, &dc);
  }
  tr.mark("patterns");
  // halo plan
  std::vector<std::vector<int32_t>> recv(n_parts), send(n_parts);
  for (int64_t i = 0; i < n_gu; ++i) recv[owner_of(gu[i])].push_back((int32_t)(n_own + i));
  for (int64_t i = 0; i < n_gp; ++i) recv[owner_of(gp[i])].push_back((int32_t)(n_own + n_gu + i));
  if (n_parts > 1) {
    std::vector<std::vector<std::vector<int32_t>>> tsend(nt, std::vector<std::vector<int32_t>>(n_parts));
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nt; ++t) {
      int64_t b, e;
      chunk_of(nc, t, b, e);
      for (int64_t i = b; i < e; ++i) {
        const int32_t *cd = &d->cell_dofs[15 * (int64_t)P->cell_ids[i]];
        int owners[15], no = 0;
        for (int k = 0; k < 15; ++k) {
          const int o = owner_of(cd[k]);
          bool seen = false;
          for (int q = 0; q < no; ++q) seen |= owners[q] == o;
          if (!seen) owners[no++] = o;
        }
        if (no == 1) continue;
        for (int q = 0; q < no; ++q) {
          if (owners[q] == rank) continue;
          for (int k = 0; k < 15; ++k)
            if (owned_g(cd[k])) tsend[t][owners[q]].push_back(g2l(cd[k]));
        }
      }
      for (auto &v : tsend[t]) uniq(v);
    }
    for (int t = 0; t < nt; ++t)
      for (int k = 0; k < n_parts; ++k) send[k].insert(send[k].end(), tsend[t][k].begin(), tsend[t][k].end());
    for (auto &s : send) uniq(s);  // local owned ids ascend with global ids inside [u | p]
  }
  P->send_ptr.push_back(0);
  P->recv_ptr.push_back(0);
  for (int k = 0; k < n_parts; ++k) {
    if (k == rank || (send[k].empty() && recv[k].empty())) continue;
    P->neighbors.push_back(k);
    P->send_idx.insert(P->send_idx.end(), send[k].begin(), send[k].end());
    P->recv_idx.insert(P->recv_idx.end(), recv[k].begin(), recv[k].end());
    P->send_ptr.push_back((int64_t)P->send_idx.size());
    P->recv_ptr.push_back((int64_t)P->recv_idx.size());
  }
  nst_part_info &I = P->info;
  I.n_own_u = n_own_u;
  I.n_own_p = n_own_p;
  I.n_ghost_u = n_gu;
  I.n_ghost_p = n_gp;
  I.n_cells = nc;
  I.n_owned_cells = 0;
  for (uint8_t o : P->cell_owned) I.n_owned_cells += o;
  I.n_vertices = nv;
  I.nnz_jac = P->jac_rowptr.empty() ? 0 : P->jac_rowptr[n_own];
  I.nnz_pm = P->pm_rowptr.empty() ? 0 : P->pm_rowptr[n_own];
  I.n_neighbors = (int32_t)P->neighbors.size();
  I.n_send = (int64_t)P->send_idx.size();
  I.n_recv = (int64_t)P->recv_idx.size();
  tr.mark("halo plan");
  *out = P;
  return NST_OK;
}

void nst_part_free(nst_part *p) { delete p; }
int nst_part_get_info(const nst_part *p, nst_part_info *info) {
  if (!p || !info) return fail(NST_ERR_ARG, "null argument");
  *info = p->info;
  return NST_OK;
}
const int64_t *nst_part_l2g(const nst_part *p) { return p->l2g.data(); }
const int32_t *nst_part_cell_ids(const nst_part *p) { return p->cell_ids.data(); }
const int32_t *nst_part_cell_dofs(const nst_part *p) { return p->cell_dofs.data(); }
const int32_t *nst_part_cell_vertices(const nst_part *p) { return p->cell_vertices.data(); }
const double *nst_part_xy(const nst_part *p) { return p->xy.data(); }
const uint8_t *nst_part_cell_owned(const nst_part *p) { return p->cell_owned.data(); }
const int64_t *nst_part_jac_rowptr(const nst_part *p) { return p->jac_rowptr.empty() ? nullptr : p->jac_rowptr.data(); }
const int32_t *nst_part_jac_col(const nst_part *p) { return p->jac_col.data(); }
const int64_t *nst_part_pm_rowptr(const nst_part *p) { return p->pm_rowptr.empty() ? nullptr : p->pm_rowptr.data(); }
const int32_t *nst_part_pm_col(const nst_part *p) { return p->pm_col.data(); }
const int32_t *nst_part_neighbors(const nst_part *p) { return p->neighbors.data(); }
const int64_t *nst_part_send_ptr(const nst_part *p) { return p->send_ptr.data(); }
const int32_t *nst_part_send_idx(const nst_part *p) { return p->send_idx.data(); }
const int64_t *nst_part_recv_ptr(const nst_part *p) { return p->recv_ptr.data(); }
const int32_t *nst_part_recv_idx(const nst_part *p) { return p->recv_idx.data(); }
int64_t nst_part_n_boundary_faces(const nst_part *p) { return (int64_t)p->bface_cell.size(); }
const int32_t *nst_part_bface_cell(const nst_part *p) { return p->bface_cell.data(); }
const int32_t *nst_part_bface_face(const nst_part *p) { return p->bface_face.data(); }
const int32_t *nst_part_bface_tag(const nst_part *p) { return p->bface_tag.data(); }

}  // extern "C"
