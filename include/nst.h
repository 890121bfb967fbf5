/* nst.h — host "topology" C-ABI (libnst.so): everything NavierStokesSolver::setup()
 * produces for the per-Newton-step hot path, restated WITHOUT deal.II
 * (deal.II / Trilinos / METIS are absent from this image: SURVEY.md §0 F1).
 *
 * It replaces, on the host and once per run, these reference calls
 * (all file:line relative to /root/reference):
 *   GridIn::read_msh                         src/NavierStokesSolver.cpp:12-16
 *   GridTools::partition_triangulation       src/NavierStokesSolver.cpp:18      (METIS -> RCB)
 *   fullydistributed::create_triangulation   src/NavierStokesSolver.cpp:19-21   (owned + ghost layer)
 *   DoFHandler::distribute_dofs              src/NavierStokesSolver.cpp:64-65
 *   DoFRenumbering::component_wise           src/NavierStokesSolver.cpp:69-73
 *   locally owned / relevant index sets      src/NavierStokesSolver.cpp:75-91
 *   DoFTools::make_sparsity_pattern x3       src/NavierStokesSolver.cpp:107-158
 *   VectorTools::interpolate_boundary_values src/NavierStokesSolver.cpp:357-373
 *
 * Plain C, opaque handles, caller-owned output buffers. No CUDA here: the
 * arrays this library hands out are what include/nsg.h uploads to the GPU.
 * All functions return 0 on success or a negative NST_ERR_* code; a text
 * message for the last error of the calling thread is in nst_last_error().
 */
#ifndef NST_H
#define NST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NST_OK 0
#define NST_ERR_IO (-1)
#define NST_ERR_FORMAT (-2)
#define NST_ERR_ARG (-3)
#define NST_ERR_TOPOLOGY (-4) /* e.g. an edge shared by 3 cells (mesh2d.msh as a whole) */

typedef struct nst_mesh nst_mesh;
typedef struct nst_dofs nst_dofs;
typedef struct nst_part nst_part;

const char *nst_last_error(void);

/* ---- mesh ------------------------------------------------------------- */

/* GridIn::read_msh restatement (gmsh ASCII 2.2 and 4.1; triangles + tagged lines).
 * surface_entity >= 0 keeps only the triangles of that gmsh surface entity (4.1)
 * or of that elementary tag (2.2); -1 keeps all. Negative-measure cells are
 * inverted by swapping vertices 1 and 2. */
int nst_mesh_read_msh(const char *path, int surface_entity, nst_mesh **out);

/* From arrays: xy[2*n_vertices], cells[3*n_cells] (0-based), tagged boundary lines
 * line_v[2*n_lines] with line_tag[n_lines] (first physical tag). */
int nst_mesh_create(int64_t n_vertices, const double *xy, int64_t n_cells, const int32_t *cells,
                    int64_t n_lines, const int32_t *line_v, const int32_t *line_tag, nst_mesh **out);

/* Uniform red refinement, `levels` times. Children of cell c are 4c..4c+3
 * (v0 m01 m20 | m01 v1 m12 | m20 m12 v2 | m01 m12 m20); boundary ids inherited.
 * If snap_id >= 0, new vertices on boundary edges with that id are projected to
 * the circle (cx,cy,r). */
int nst_mesh_refine(const nst_mesh *in, int levels, int snap_id, double cx, double cy, double r,
                    nst_mesh **out);

/* Geometric boundary ids for untagged meshes (mesh2d.msh, mesh_poli*.msh): boundary
 * edges whose midpoint has x==xmin -> id_left, x==xmax -> id_right, y==ymin/ymax ->
 * id_wall, everything else -> id_other (bounding box computed from the mesh). */
int nst_mesh_tag_boundary_box(nst_mesh *m, int id_left, int id_right, int id_wall, int id_other);

void nst_mesh_free(nst_mesh *m);

int64_t nst_mesh_n_vertices(const nst_mesh *m);
int64_t nst_mesh_n_cells(const nst_mesh *m);
int64_t nst_mesh_n_edges(const nst_mesh *m);
int64_t nst_mesh_n_boundary_edges(const nst_mesh *m);
int64_t nst_mesh_n_inverted(const nst_mesh *m); /* cells whose orientation was fixed */
const double *nst_mesh_xy(const nst_mesh *m);          /* [2V] */
const int32_t *nst_mesh_cells(const nst_mesh *m);      /* [3T] */
const int32_t *nst_mesh_cell_edges(const nst_mesh *m); /* [3T] line l of a cell: 0=(v0,v1) 1=(v1,v2) 2=(v2,v0) */
const int32_t *nst_mesh_edge_vertices(const nst_mesh *m); /* [2E] */
const int32_t *nst_mesh_edge_tag(const nst_mesh *m);      /* [E] boundary id, -1 for interior edges */
/* boundary edges in (cell, local face) order: out_cell[nb], out_face[nb], out_tag[nb] */
int nst_mesh_boundary_faces(const nst_mesh *m, int32_t *out_cell, int32_t *out_face, int32_t *out_tag);

/* ---- partition (replaces METIS k-way by recursive coordinate bisection) -------- */
int nst_partition_rcb(const nst_mesh *m, int n_parts, int32_t *cell_part /* [T] */);

/* ---- DoFs: FESystem(FE_SimplexP(2)^2, FE_SimplexP(1)), 15 per cell ----------- */
/* cell_part == NULL or n_parts == 1: the serial numbering. Otherwise the deal.II
 * parallel numbering: DoFs on a subdomain interface belong to the lowest part id;
 * block-wise, part-major inside each block. Local cell order:
 * i = 3v+{0,1} (u_x,u_y) and 3v+2 (p) for vertex v; 9+2l+{0,1} for line l. */
int nst_dofs_distribute(const nst_mesh *m, int n_parts, const int32_t *cell_part, nst_dofs **out);
void nst_dofs_free(nst_dofs *d);
int64_t nst_dofs_n_u(const nst_dofs *d);
int64_t nst_dofs_n_p(const nst_dofs *d);
const int32_t *nst_dofs_cell_dofs(const nst_dofs *d);  /* [15T] global indices */
const int32_t *nst_dofs_vertex_node(const nst_dofs *d); /* [V]  P2-node id of a vertex: u dofs 2n,2n+1 */
const int32_t *nst_dofs_edge_node(const nst_dofs *d);   /* [E]  P2-node id of an edge midpoint */
const int32_t *nst_dofs_vertex_p(const nst_dofs *d);    /* [V]  pressure dof index - n_u */
/* per-part owned counts (n_parts entries each) */
const int64_t *nst_dofs_part_n_u(const nst_dofs *d);
const int64_t *nst_dofs_part_n_p(const nst_dofs *d);

/* Sparsity patterns as CSR over the global numbering, columns ascending.
 * kind: 0 = Jacobian, full coupling incl. p-p (cpp:107-110); 1 = Stokes, no p-p
 * (cpp:124-140); 2 = pressure mass, p-p only (cpp:143-158; rows < n_u are empty).
 * Call once with rowptr/col NULL to get nnz, then with buffers. */
int nst_sparsity(const nst_mesh *m, const nst_dofs *d, int kind, int64_t *nnz,
                 int64_t *rowptr /* [N+1] */, int32_t *col /* [nnz] */);

/* interpolate_boundary_values restatement. `calls` lists boundary ids in groups:
 * ids[] / is_inlet[] of length n_ids, call_ptr[n_calls+1] delimits the successive
 * interpolate_boundary_values calls that accumulate into one ordered map
 * (cpp:357-373: call 0 = {11:inlet}, call 1 = {11:inlet,12:zero,13:zero}).
 * Velocity components only (ComponentMask {true,true,false}).
 * Inlet: u_x = 4 u_m (y-y0)(H-(y-y0))/H^2 * time_factor, u_y = 0 (hpp:457).
 * Output sorted by dof; returns the count via *n_out (call with NULL buffers first). */
typedef struct {
  double u_m, H, y0, time_factor;
} nst_inlet_params;
int nst_dirichlet_values(const nst_mesh *m, const nst_dofs *d, int n_calls, const int32_t *call_ptr,
                         const int32_t *ids, const int32_t *is_inlet, const nst_inlet_params *inlet,
                         int64_t *n_out, int32_t *out_dof, double *out_val);

/* Support point of every global DoF: xy[2N] (vertices and edge midpoints). */
int nst_dofs_support_points(const nst_mesh *m, const nst_dofs *d, double *xy);

/* ---- one rank's local problem (owned rows + ghost layer) -------------------- */
/* Local numbering: [owned u | owned p | ghost u | ghost p]; owned in global order,
 * ghosts sorted by (owner, global id). Local cells = every cell touching an owned
 * DoF (owned cells + the part of the ghost layer that is needed), in global order. */
int nst_part_build(const nst_mesh *m, const nst_dofs *d, int n_parts, const int32_t *cell_part,
                   int rank, nst_part **out);
/* The same with flags: NST_PART_NO_PATTERNS leaves the two sparsity patterns (cpp:101-158) out - the device builds them
 * from the cell -> dof table (nsg_set_pattern_from_cells); rowptrs are all zero, nnz_jac = nnz_pm = 0. */
#define NST_PART_NO_PATTERNS 1
int nst_part_build_ex(const nst_mesh *m, const nst_dofs *d, int n_parts, const int32_t *cell_part,
                      int rank, int flags, nst_part **out);
void nst_part_free(nst_part *p);
typedef struct {
  int64_t n_own_u, n_own_p, n_ghost_u, n_ghost_p;
  int64_t n_cells, n_owned_cells, n_vertices;
  int64_t nnz_jac, nnz_pm;
  int32_t n_neighbors;
  int64_t n_send, n_recv;
} nst_part_info;
int nst_part_get_info(const nst_part *p, nst_part_info *info);
const int64_t *nst_part_l2g(const nst_part *p);          /* [n_loc] local -> global dof */
const int32_t *nst_part_cell_ids(const nst_part *p);     /* [n_cells] global cell id */
const int32_t *nst_part_cell_dofs(const nst_part *p);    /* [15 n_cells] local dof ids */
const int32_t *nst_part_cell_vertices(const nst_part *p);/* [3 n_cells] local vertex ids */
const double *nst_part_xy(const nst_part *p);            /* [2 n_vertices] */
const uint8_t *nst_part_cell_owned(const nst_part *p);   /* [n_cells] 1 if the cell is owned by rank */
const int64_t *nst_part_jac_rowptr(const nst_part *p);   /* [n_own+1]; NULL for a part built with NST_PART_NO_PATTERNS */
const int32_t *nst_part_jac_col(const nst_part *p);      /* local column ids, ascending */
const int64_t *nst_part_pm_rowptr(const nst_part *p);    /* [n_own+1], rows < n_own_u empty; NULL with NST_PART_NO_PATTERNS */
const int32_t *nst_part_pm_col(const nst_part *p);
/* halo plan: neighbours in ascending rank; send_idx are local owned indices to pack for
 * neighbour k in [send_ptr[k],send_ptr[k+1]); recv for neighbour k lands contiguously at
 * local indices recv_idx[recv_ptr[k] .. recv_ptr[k+1]). */
const int32_t *nst_part_neighbors(const nst_part *p);
const int64_t *nst_part_send_ptr(const nst_part *p);
const int32_t *nst_part_send_idx(const nst_part *p);
const int64_t *nst_part_recv_ptr(const nst_part *p);
const int32_t *nst_part_recv_idx(const nst_part *p);
/* boundary faces of local cells: (local cell, face, tag) */
int64_t nst_part_n_boundary_faces(const nst_part *p);
const int32_t *nst_part_bface_cell(const nst_part *p);
const int32_t *nst_part_bface_face(const nst_part *p);
const int32_t *nst_part_bface_tag(const nst_part *p);

#ifdef __cplusplus
}
#endif
#endif
