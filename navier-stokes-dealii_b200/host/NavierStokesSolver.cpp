// NavierStokesSolver.cpp — see NavierStokesSolver.hpp.  Control flow, tolerances and printed
// quantities follow /root/reference src/NavierStokesSolver.cpp; all array work is behind
// include/nst.h (host topology) and include/nsg.h (CUDA).  Non-zero return codes of the C ABI become
// exceptions, mirroring the reference's uncaught deal.II exceptions (SURVEY §8b "Conventions").
#include "NavierStokesSolver.hpp"

#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <thread>

#include "nsg.h"
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "nst.h"

namespace {
int env_int(const char *a, const char *b, int dflt) {
  for (const char *n : {a, b})
    if (n)
      if (const char *v = std::getenv(n)) return std::atoi(v);
  return dflt;
}
double env_double(const char *n, double dflt) {
  const char *v = std::getenv(n);
  return v ? std::atof(v) : dflt;
}
std::string env_str(const char *n, const std::string &dflt) {
  const char *v = std::getenv(n);
  return v ? std::string(v) : dflt;
}
// what all ranks of one launch share and a later launch does not: NS_RENDEZVOUS_NONCE, else the parent process
// (torchrun agent / mpirun) identified by pid and start time (field 22 of /proc/<pid>/stat)
std::string launch_nonce() {
  if (const char *v = std::getenv("NS_RENDEZVOUS_NONCE")) return v;
  const long ppid = (long)::getppid();
  std::string start = "0";
  std::ifstream f("/proc/" + std::to_string(ppid) + "/stat");
  std::string line;
  if (f && std::getline(f, line)) {
    const size_t rp = line.rfind(')');   // the command name may contain spaces
    std::stringstream ss(rp == std::string::npos ? line : line.substr(rp + 1));
    std::string tok;
    for (int i = 0; i < 20 && (ss >> tok); ++i) start = tok;   // 20th field after the name = starttime
  }
  return std::to_string(ppid) + "-" + start;
}
std::vector<int> env_list(const char *n, std::vector<int> dflt) {
  const char *v = std::getenv(n);
  if (!v) return dflt;
  std::vector<int> out;
  std::stringstream ss(v);
  std::string tok;
  while (std::getline(ss, tok, ',')) out.push_back(std::atoi(tok.c_str()));
  return out;
}
}  // namespace

namespace ns_b200_compat {
namespace Utilities {
namespace MPI {
MPI_InitFinalize::MPI_InitFinalize(int &, char **&) {}
MPI_InitFinalize::~MPI_InitFinalize() {}
unsigned int n_mpi_processes() { return (unsigned)env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", 1); }
unsigned int this_mpi_process() { return (unsigned)env_int("RANK", "OMPI_COMM_WORLD_RANK", 0); }
}  // namespace MPI
}  // namespace Utilities
}  // namespace ns_b200_compat

#define NST_CALL(expr)                                                        \
  do {                                                                        \
    if ((expr) != NST_OK) fail(std::string(#expr) + ": " + nst_last_error()); \
  } while (0)
#define NSG_CALL(expr)                                                        \
  do {                                                                        \
    if ((expr) != NSG_OK) fail(std::string(#expr) + ": " + nsg_last_error()); \
  } while (0)

NavierStokesSolver::NavierStokesSolver(const unsigned int &degree_velocity_, const unsigned int &degree_pressure_, const double &T_,
                                       const double &deltat_)
    : mpi_size(Utilities::MPI::n_mpi_processes()),
      mpi_rank(Utilities::MPI::this_mpi_process()),
      pcout(std::cout, mpi_rank == 0),
      T(T_),
      degree_velocity(degree_velocity_),
      degree_pressure(degree_pressure_),
      deltat(deltat_) {
  mesh_path = env_str("NS_MESH", "../mesh/correct_mesh_yt.msh");  // cpp:15
  surface_entity = env_int("NS_SURFACE_ENTITY", nullptr, -1);
  refine_levels = env_int("NS_REFINE", nullptr, 0);
  nu = env_double("NS_NU", nu), rho = env_double("NS_RHO", rho), p_out = env_double("NS_P_OUT", p_out);
  g = env_double("NS_G", g), u_m = env_double("NS_U_M", u_m), H = env_double("NS_H", H);
  inlet_y0 = env_double("NS_INLET_Y0", inlet_y0);
  inlet_time_mode = env_str("NS_INLET_TIME", "frozen");  // set_time() is never called in the reference
  neumann_id = env_int("NS_NEUMANN_ID", nullptr, 10);    // cpp:320
  inlet_id = env_int("NS_INLET_ID", nullptr, 11);        // cpp:357
  wall_ids = env_list("NS_WALL_IDS", {12, 13});          // cpp:367-368
  box_tags = env_list("NS_BOX_TAGS", {});
  preconditioner = env_str("NS_PRECONDITIONER", "identity");  // cpp:570
  stokes_init = env_int("NS_STOKES_INIT", nullptr, 0) != 0;   // cpp:636-644 is commented out
  output_dir = env_str("NS_OUTPUT_DIR", "");
}

NavierStokesSolver::~NavierStokesSolver() {
  if (dev) nsg_destroy(dev);
  if (part) nst_part_free(part);
  if (dofs) nst_dofs_free(dofs);
  if (mesh) nst_mesh_free(mesh);
}

void NavierStokesSolver::fail(const std::string &what) const { throw std::runtime_error("NavierStokesSolver: " + what); }

void NavierStokesSolver::setup() {
  if (degree_velocity != 2 || degree_pressure != 1) fail("the B200 path implements the P2-P1 pair of main.cpp:9-10");
  pcout << "Initializing the mesh" << std::endl;
  NST_CALL(nst_mesh_read_msh(mesh_path.c_str(), surface_entity, &mesh));
  if (box_tags.size() == 4) NST_CALL(nst_mesh_tag_boundary_box(mesh, box_tags[0], box_tags[1], box_tags[2], box_tags[3]));
  if (refine_levels > 0) {
    nst_mesh *fine = nullptr;
    NST_CALL(nst_mesh_refine(mesh, refine_levels, -1, 0, 0, 0, &fine));
    nst_mesh_free(mesh);
    mesh = fine;
  }
  pcout << "  Number of elements = " << nst_mesh_n_cells(mesh) << std::endl;
  pcout << "-----------------------------------------------" << std::endl;
  pcout << "Initializing the finite element space" << std::endl;
  pcout << "  Velocity degree:           = " << degree_velocity << std::endl;
  pcout << "  Pressure degree:           = " << degree_pressure << std::endl;
  pcout << "  DoFs per cell              = " << 15 << std::endl;
  pcout << "  Quadrature points per cell = " << 7 << std::endl;
  pcout << "  Quadrature points per face = " << 3 << std::endl;
  pcout << "-----------------------------------------------" << std::endl;
  pcout << "Initializing the DoF handler" << std::endl;
  std::vector<int32_t> cell_part;
  if (mpi_size > 1) {
    cell_part.resize(nst_mesh_n_cells(mesh));
    NST_CALL(nst_partition_rcb(mesh, (int)mpi_size, cell_part.data()));
  }
  NST_CALL(nst_dofs_distribute(mesh, (int)mpi_size, mpi_size > 1 ? cell_part.data() : nullptr, &dofs));
  n_u_global = nst_dofs_n_u(dofs), n_p_global = nst_dofs_n_p(dofs);
  pcout << "  Number of DoFs: " << std::endl;
  pcout << "    velocity = " << n_u_global << std::endl;
  pcout << "    pressure = " << n_p_global << std::endl;
  pcout << "    total    = " << n_u_global + n_p_global << std::endl;
  pcout << "-----------------------------------------------" << std::endl;
  pcout << "  Initializing the linear system" << std::endl;
  // the three sparsity patterns of cpp:101-158 are built on the device from the cell -> dof table;
  // NS_HOST_PATTERNS=1 builds them in libnst and uploads them instead (bit-identical, slower)
  const bool host_patterns = env_int("NS_HOST_PATTERNS", nullptr, 0) != 0;
  NST_CALL(nst_part_build_ex(mesh, dofs, (int)mpi_size, mpi_size > 1 ? cell_part.data() : nullptr, (int)mpi_rank,
                             host_patterns ? 0 : NST_PART_NO_PATTERNS, &part));
  nst_part_info I;
  NST_CALL(nst_part_get_info(part, &I));
  n_own_u = I.n_own_u, n_own = I.n_own_u + I.n_own_p;
  for (unsigned r = 0; r < mpi_rank; ++r) u_lo += nst_dofs_part_n_u(dofs)[r], p_lo += nst_dofs_part_n_p(dofs)[r];
  const int device = env_int("LOCAL_RANK", "OMPI_COMM_WORLD_LOCAL_RANK", 0);
  NSG_CALL(nsg_create(device, &dev));
  pcout << "  Initializing the sparsity pattern" << std::endl;
  if (host_patterns)
    NSG_CALL(nsg_set_pattern(dev, I.n_own_u, I.n_own_p, I.n_ghost_u, I.n_ghost_p, nst_part_jac_rowptr(part), nst_part_jac_col(part),
                             nst_part_pm_rowptr(part), nst_part_pm_col(part)));
  else
    NSG_CALL(nsg_set_pattern_from_cells(dev, I.n_own_u, I.n_own_p, I.n_ghost_u, I.n_ghost_p, I.n_cells, nst_part_cell_dofs(part)));
  pcout << "  Initializing the matrices" << std::endl;
  NSG_CALL(nsg_set_mesh(dev, I.n_cells, I.n_vertices, nst_part_xy(part), nst_part_cell_vertices(part), nst_part_cell_dofs(part),
                        nst_part_n_boundary_faces(part), nst_part_bface_cell(part), nst_part_bface_face(part),
                        nst_part_bface_tag(part)));
  if (mpi_size > 1) {
    // ncclUniqueId from rank 0 through a rendezvous file (no MPI in this image).  The file belongs to ONE launch: its
    // name and its header carry a nonce all ranks of the launch share (the launcher's pid and start time - the ranks
    // are siblings under torchrun / mpirun - or NS_RENDEZVOUS_NONCE), rank 0 removes a stale file before it creates the
    // new one with O_EXCL | O_NOFOLLOW (mode 0600), readers accept only a regular file of their own uid whose header
    // matches, and rank 0 removes the file once the communicator is up.
    const std::string nonce = launch_nonce();
    const std::string path = env_str("NS_RENDEZVOUS", "/tmp/ns_nccl_id." + env_str("MASTER_PORT", "0")) + "." + nonce;
    char id[128];
    char header[64];
    std::memset(header, 0, sizeof header);
    std::snprintf(header, sizeof header, "NSNCCLID1 %s", nonce.c_str());
    if (mpi_rank == 0) {
      NSG_CALL(nsg_comm_unique_id(id));
      ::unlink(path.c_str());
      const std::string tmp = path + ".tmp";
      ::unlink(tmp.c_str());
      const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
      if (fd < 0) fail("cannot create the NCCL id file " + tmp);
      const bool ok = ::write(fd, header, sizeof header) == (ssize_t)sizeof header && ::write(fd, id, 128) == 128;
      ::close(fd);
      if (!ok || std::rename(tmp.c_str(), path.c_str()) != 0) fail("cannot write the NCCL id file " + path);
    } else {
      for (int tries = 0;; ++tries) {
        const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
        if (fd >= 0) {
          struct stat st;
          char h[64];
          const bool ok = ::fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == ::getuid() &&
                          ::read(fd, h, sizeof h) == (ssize_t)sizeof h && std::memcmp(h, header, sizeof h) == 0 &&
                          ::read(fd, id, 128) == 128;
          ::close(fd);
          if (ok) break;
        }
        if (tries > 6000) fail("timed out waiting for the NCCL id file " + path);
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
      }
    }
    NSG_CALL(nsg_comm_init(dev, (int)mpi_rank, (int)mpi_size, id));  // collective: every rank has read the file when it returns
    if (mpi_rank == 0) ::unlink(path.c_str());
    NSG_CALL(nsg_set_halo(dev, I.n_neighbors, nst_part_neighbors(part), nst_part_send_ptr(part), nst_part_send_idx(part),
                          nst_part_recv_ptr(part), nst_part_recv_idx(part)));
  }
  pcout << "  Initializing the system right-hand side" << std::endl;
  pcout << "  Initializing the solution vector" << std::endl;
  push_params(false);
}

void NavierStokesSolver::push_params(bool stokes) {
  nsg_params P;
  nsg_params_default(&P);
  P.nu = nu, P.rho = rho, P.p_out = p_out, P.deltat = deltat;
  P.forcing[0] = 0.0, P.forcing[1] = -g;
  P.neumann_id = stokes ? env_int("NS_STOKES_NEUMANN_ID", nullptr, 1) : neumann_id;
  P.use_mass = env_int("NS_USE_MASS", nullptr, 1);
  P.stokes = stokes ? 1 : 0;
  P.dirichlet_diag = env_int("NS_DIRICHLET_DIAG", nullptr, 0);  // 0: TrilinosWrappers rule (reference path), 1: keep a non-zero diagonal
  NSG_CALL(nsg_set_params(dev, &P));
}

void NavierStokesSolver::dirichlet(bool stokes, std::vector<int32_t> &ldofs, std::vector<double> &lvals) const {
  // interpolate_boundary_values twice into one map (cpp:351-373); boundary_functions.clear() is commented
  // out for the Navier-Stokes path (cpp:364) and active for the Stokes path (cpp:516)
  std::vector<int32_t> ids, is_inlet, call_ptr{0};
  const int in_id = stokes ? env_int("NS_STOKES_INLET_ID", nullptr, 0) : inlet_id;
  const std::vector<int> walls = stokes ? env_list("NS_STOKES_WALL_IDS", {2, 3}) : wall_ids;
  ids.push_back(in_id), is_inlet.push_back(1);
  call_ptr.push_back((int32_t)ids.size());
  if (!stokes) ids.push_back(in_id), is_inlet.push_back(1);
  for (int w : walls) ids.push_back(w), is_inlet.push_back(0);
  call_ptr.push_back((int32_t)ids.size());
  nst_inlet_params ip;
  ip.u_m = u_m, ip.H = H, ip.y0 = inlet_y0;
  const double t = inlet_time_mode == "frozen" ? 0.0 : time;
  ip.time_factor = inlet_time_mode == "constant" ? 1.0 : std::sin(M_PI * t / 8.);  // hpp:457
  int64_t n = 0;
  NST_CALL(nst_dirichlet_values(mesh, dofs, 2, call_ptr.data(), ids.data(), is_inlet.data(), &ip, &n, nullptr, nullptr));
  std::vector<int32_t> gd(n > 0 ? n : 1);
  std::vector<double> gv(n > 0 ? n : 1);
  NST_CALL(nst_dirichlet_values(mesh, dofs, 2, call_ptr.data(), ids.data(), is_inlet.data(), &ip, &n, gd.data(), gv.data()));
  ldofs.clear(), lvals.clear();
  // NS_INCREMENT_BC=consistent imposes g - u^k on the Newton increment; the default reproduces the reference,
  // which imposes the full g on every increment (cpp:375-376; harmless as shipped because g == 0)
  std::vector<double> cur;
  if (!stokes && env_str("NS_INCREMENT_BC", "reference") == "consistent") cur = solution_owned_values();
  const int64_t nup = n_own - n_own_u;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t gdof = gd[i];
    if (gdof < n_u_global) {
      if (gdof >= u_lo && gdof < u_lo + n_own_u) {
        ldofs.push_back((int32_t)(gdof - u_lo));
        lvals.push_back(cur.empty() ? gv[i] : gv[i] - cur[gdof - u_lo]);
      }
    } else if (gdof - n_u_global >= p_lo && gdof - n_u_global < p_lo + nup)
      ldofs.push_back((int32_t)(n_own_u + gdof - n_u_global - p_lo)), lvals.push_back(gv[i]);
  }
}

void NavierStokesSolver::assemble_system() {
  pcout << "===============================================" << std::endl;
  pcout << "Assembling the system" << std::endl;
  NSG_CALL(nsg_assemble(dev));
  std::vector<int32_t> d;
  std::vector<double> v;
  dirichlet(false, d, v);
  NSG_CALL(nsg_apply_dirichlet(dev, (int64_t)d.size(), d.data(), v.data(), 0));
}

void NavierStokesSolver::solve_system() {
  pcout << "===============================================" << std::endl;
  const int kind = preconditioner == "block_diagonal" ? NSG_PRECOND_BLOCK_DIAGONAL
                   : preconditioner == "block_triangular" ? NSG_PRECOND_BLOCK_TRIANGULAR
                                                          : NSG_PRECOND_IDENTITY;
  pcout << "Solving system..." << std::endl;
  int32_t its = 0;
  double res = 0;
  NSG_CALL(nsg_solve(dev, kind, 1e-2, 100000, 30, 0, &its, &res));  // SolverControl(100000, 1e-2 * ||R||), cpp:566
  pcout << "   " << its << " GMRES iterations" << std::endl;
  history_.back().gmres_steps = its;
}

void NavierStokesSolver::assemble_stokes_system() {
  pcout << "===============================================" << std::endl;
  pcout << "Assembling the Stokes system" << std::endl;
  push_params(true);
  NSG_CALL(nsg_assemble(dev));
  std::vector<int32_t> d;
  std::vector<double> v;
  dirichlet(true, d, v);
  NSG_CALL(nsg_apply_dirichlet(dev, (int64_t)d.size(), d.data(), v.data(), 1));
  push_params(false);
}

void NavierStokesSolver::solve_stokes_system() {
  pcout << "===============================================" << std::endl;
  pcout << "Solving the Stokes system" << std::endl;
  int32_t its = 0;
  double res = 0;
  NSG_CALL(nsg_solve(dev, NSG_PRECOND_BLOCK_TRIANGULAR, 1e-6, 2000, 30, 1, &its, &res));  // cpp:537-552
  pcout << "  " << its << " GMRES iterations" << std::endl;
  output(0., 0.);
}

void NavierStokesSolver::solve_newton() {
  const unsigned int n_max_iters = 1000;     // cpp:593
  const double residual_tolerance = 1e-2;    // cpp:594
  unsigned int n_iter = 0;
  double residual_norm = residual_tolerance + 1;
  while (n_iter < n_max_iters && residual_norm > residual_tolerance) {
    assemble_system();
    NSG_CALL(nsg_residual_norm(dev, &residual_norm));
    pcout << "  Newton iteration " << n_iter << "/" << n_max_iters << " - ||r|| = " << std::scientific << std::setprecision(6)
          << residual_norm << std::flush;
    history_.push_back({current_step, n_iter, residual_norm, -1});
    if (residual_norm > residual_tolerance) {
      solve_system();
      pcout << "System solved!" << std::endl;
      NSG_CALL(nsg_update_solution(dev));  // solution_owned += delta_owned; solution = solution_owned
    } else {
      pcout << " < tolerance" << std::endl;
    }
    ++n_iter;
  }
}

void NavierStokesSolver::solve() {
  pcout << "===============================================" << std::endl;
  time = 0.0;
  if (stokes_init) {
    pcout << "Finding the initial condition" << std::endl;
    assemble_stokes_system();
    solve_stokes_system();
    pcout << "-----------------------------------------------" << std::endl;
  } else {
    pcout << "Applying the initial condition" << std::endl;
    std::vector<double> zero(n_own > 0 ? n_own : 1, 0.0);  // FunctionU0 == 0 (hpp:478-497)
    NSG_CALL(nsg_set_solution(dev, zero.data()));
    output(0, 0.0);
    pcout << "-----------------------------------------------" << std::endl;
  }
  unsigned int time_step = 0;
  while (time < T - 0.5 * deltat) {
    time += deltat;
    ++time_step;
    current_step = time_step;
    NSG_CALL(nsg_push_time_level(dev));  // solution_old = solution
    pcout << "n = " << std::setw(3) << time_step << ", t = " << std::setw(5) << std::fixed << time << std::endl;
    solve_newton();
    output(time_step, time);
    pcout << std::endl;
  }
}

std::vector<double> NavierStokesSolver::solution_owned_values() const {
  std::vector<double> v(n_own > 0 ? n_own : 1);
  if (nsg_get_solution(dev, v.data()) != NSG_OK) fail(nsg_last_error());
  v.resize(n_own);
  return v;
}

// Minimal writer (HDF5/XDMF of the reference need libraries absent here): legacy VTK, P1 view of the
// owned+ghost cells of this rank with velocity/pressure at the vertices and the partition id per cell.
void NavierStokesSolver::output(const unsigned int &time_step, const double &t) const {
  pcout << "===============================================" << std::endl;
  if (output_dir.empty()) return;
  nst_part_info I;
  nst_part_get_info(part, &I);
  // the ghosted `solution` DataOut reads (cpp:697-700): owned entries + ghost layer
  std::vector<double> sol((size_t)std::max<int64_t>(I.n_own_u + I.n_own_p + I.n_ghost_u + I.n_ghost_p, 1));
  NSG_CALL(nsg_get_solution_ghosted(dev, sol.data()));
  const int32_t *cv = nst_part_cell_vertices(part), *cd = nst_part_cell_dofs(part);
  const double *xy = nst_part_xy(part);
  const uint8_t *owned = nst_part_cell_owned(part);
  std::vector<int64_t> cells;
  for (int64_t c = 0; c < I.n_cells; ++c)
    if (owned[c]) cells.push_back(c);
  const size_t T = cells.size();
  std::ostringstream stem;
  stem << "output-" << std::setw(4) << std::setfill('0') << time_step;
  // one patch of three nodes per owned cell, as DataOut::build_patches with filter_duplicate_vertices = false
  // (cpp:685-719): points, triangles, velocity (3-component), pressure, partitioning
  std::vector<double> pts(6 * T), vel(9 * T, 0.0), pres(3 * T), partn(3 * T, (double)mpi_rank);
  std::vector<int32_t> tri(3 * T);
  for (size_t i = 0; i < T; ++i) {
    const int64_t c = cells[i];
    for (int k = 0; k < 3; ++k) {
      pts[6 * i + 2 * k] = xy[2 * cv[3 * c + k]], pts[6 * i + 2 * k + 1] = xy[2 * cv[3 * c + k] + 1];
      vel[9 * i + 3 * k] = sol[cd[15 * c + 3 * k]], vel[9 * i + 3 * k + 1] = sol[cd[15 * c + 3 * k + 1]];
      pres[3 * i + k] = sol[cd[15 * c + 3 * k + 2]];
      tri[3 * i + k] = (int32_t)(3 * i + k);
    }
  }
  // (1) XDMF + raw little-endian heavy data (HDF5 is not available: Format="Binary" with Seek offsets); the layout is
  //     the one navier-stokes-dealii_b200/output.py writes and reads back
  {
    const std::string bin = stem.str() + ".rank" + std::to_string(mpi_rank) + ".bin";
    std::ofstream fb(output_dir + "/" + bin, std::ios::binary);
    if (!fb) fail("cannot write " + output_dir + "/" + bin);
    size_t off[5], o = 0;
    auto put = [&](const void *ptr, size_t bytes, int slot) {
      off[slot] = o;
      fb.write((const char *)ptr, (std::streamsize)bytes);
      o += bytes;
    };
    put(pts.data(), 8 * pts.size(), 0), put(tri.data(), 4 * tri.size(), 1), put(vel.data(), 8 * vel.size(), 2);
    put(pres.data(), 8 * pres.size(), 3), put(partn.data(), 8 * partn.size(), 4);
    const std::string xname = output_dir + "/" + stem.str() + (mpi_size > 1 ? ".rank" + std::to_string(mpi_rank) : std::string()) + ".xdmf";
    std::ofstream fx(xname);
    if (!fx) fail("cannot write " + xname);
    auto item = [&](const std::string &dims, const char *type, int prec, int slot) {
      std::ostringstream q;
      q << "<DataItem Dimensions=\"" << dims << "\" NumberType=\"" << type << "\" Precision=\"" << prec
        << "\" Format=\"Binary\" Endian=\"Little\" Seek=\"" << off[slot] << "\">" << bin << "</DataItem>";
      return q.str();
    };
    const std::string n3 = std::to_string(3 * T);
    fx << "<?xml version=\"1.0\" ?>\n<!DOCTYPE Xdmf SYSTEM \"Xdmf.dtd\" []>\n<Xdmf Version=\"3.0\">\n <Domain>\n"
       << "  <Grid Name=\"CellTime\" GridType=\"Collection\" CollectionType=\"Temporal\">\n"
       << "   <Grid Name=\"mesh\" GridType=\"Collection\" CollectionType=\"Spatial\">\n    <Time Value=\"" << std::setprecision(17) << t
       << "\"/>\n    <Grid Name=\"rank" << mpi_rank << "\" GridType=\"Uniform\">\n"
       << "     <Topology TopologyType=\"Triangle\" NumberOfElements=\"" << T << "\">" << item(std::to_string(T) + " 3", "Int", 4, 1)
       << "</Topology>\n     <Geometry GeometryType=\"XY\">" << item(n3 + " 2", "Float", 8, 0) << "</Geometry>\n"
       << "     <Attribute Name=\"velocity\" AttributeType=\"Vector\" Center=\"Node\">" << item(n3 + " 3", "Float", 8, 2) << "</Attribute>\n"
       << "     <Attribute Name=\"pressure\" AttributeType=\"Scalar\" Center=\"Node\">" << item(n3, "Float", 8, 3) << "</Attribute>\n"
       << "     <Attribute Name=\"partitioning\" AttributeType=\"Scalar\" Center=\"Node\">" << item(n3, "Float", 8, 4) << "</Attribute>\n"
       << "    </Grid>\n   </Grid>\n  </Grid>\n </Domain>\n</Xdmf>\n";
  }
  // (2) the same patches as legacy VTK (one self-contained ASCII file per rank and step)
  std::ofstream f(output_dir + "/" + stem.str() + ".rank" + std::to_string(mpi_rank) + ".vtk");
  if (!f) fail("cannot write the .vtk file in " + output_dir);
  f << "# vtk DataFile Version 3.0\nNavier-Stokes t=" << t << "\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS " << 3 * T << " double\n";
  for (size_t i = 0; i < 3 * T; ++i) f << pts[2 * i] << " " << pts[2 * i + 1] << " 0\n";
  f << "CELLS " << T << " " << 4 * T << "\n";
  for (size_t i = 0; i < T; ++i) f << "3 " << 3 * i << " " << 3 * i + 1 << " " << 3 * i + 2 << "\n";
  f << "CELL_TYPES " << T << "\n";
  for (size_t i = 0; i < T; ++i) f << "5\n";
  f << "POINT_DATA " << 3 * T << "\nVECTORS velocity double\n";
  for (size_t i = 0; i < 3 * T; ++i) f << vel[3 * i] << " " << vel[3 * i + 1] << " 0\n";
  f << "SCALARS pressure double 1\nLOOKUP_TABLE default\n";
  for (size_t i = 0; i < 3 * T; ++i) f << pres[i] << "\n";
  f << "CELL_DATA " << T << "\nSCALARS partitioning int 1\nLOOKUP_TABLE default\n";
  for (size_t i = 0; i < T; ++i) f << mpi_rank << "\n";
}
