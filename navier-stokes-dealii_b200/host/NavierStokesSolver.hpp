// NavierStokesSolver.hpp — C++ host shim with the reference class interface
// (/root/reference src/NavierStokesSolver.hpp:407-795) over the two C-ABI libraries
// (include/nst.h host topology, include/nsg.h CUDA hot path).  src/main.cpp of the reference
// compiles against this header unchanged: same constructor, setup(), solve(), and the
// `Utilities::MPI::MPI_InitFinalize` object it creates first.
//
// Every constant the reference hard-codes is a run-time parameter here, read from environment
// variables whose defaults are the reference's values (SURVEY §5 "config / flags"), so main.cpp stays
// untouched:  NS_MESH, NS_SURFACE_ENTITY, NS_REFINE, NS_NU, NS_RHO, NS_P_OUT, NS_G, NS_U_M, NS_H,
// NS_INLET_Y0, NS_INLET_TIME (frozen|live|constant), NS_NEUMANN_ID, NS_INLET_ID, NS_WALL_IDS ("12,13"),
// NS_PRECONDITIONER (identity|block_diagonal|block_triangular), NS_STOKES_INIT, NS_OUTPUT_DIR,
// NS_BOX_TAGS ("left,right,wall,other" geometric ids for untagged meshes), NS_INCREMENT_BC (reference|consistent).
// Ranks: RANK / WORLD_SIZE / LOCAL_RANK (torchrun style) or OMPI_COMM_WORLD_*; the NCCL id travels
// through the file NS_RENDEZVOUS (default /tmp/ns_nccl_id.<MASTER_PORT>).
#ifndef NAVIER_STOKES_SOLVER_B200_HPP
#define NAVIER_STOKES_SOLVER_B200_HPP

#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

struct nst_mesh;
struct nst_dofs;
struct nst_part;
struct nsg_ctx;

// the two deal.II names main.cpp uses besides the class
namespace ns_b200_compat {
namespace Utilities {
namespace MPI {
class MPI_InitFinalize {
public:
  MPI_InitFinalize(int &argc, char **&argv);
  ~MPI_InitFinalize();
};
unsigned int n_mpi_processes();
unsigned int this_mpi_process();
}  // namespace MPI
}  // namespace Utilities
}  // namespace ns_b200_compat
using namespace ns_b200_compat;

// rank-0-only stream (ConditionalOStream, hpp:698)
class ConditionalOStream {
public:
  ConditionalOStream(std::ostream &s, bool active) : out(s), on(active) {}
  template <class T>
  const ConditionalOStream &operator<<(const T &v) const {
    if (on) out << v;
    return *this;
  }
  const ConditionalOStream &operator<<(std::ostream &(*m)(std::ostream &)) const {
    if (on) out << m;
    return *this;
  }

private:
  std::ostream &out;
  bool on;
};

class NavierStokesSolver {
public:
  static constexpr unsigned int dim = 2;  // hpp:411

  NavierStokesSolver(const unsigned int &degree_velocity_, const unsigned int &degree_pressure_, const double &T_,
                     const double &deltat_);
  ~NavierStokesSolver();
  NavierStokesSolver(const NavierStokesSolver &) = delete;
  NavierStokesSolver &operator=(const NavierStokesSolver &) = delete;

  void setup();  // cpp:4-176
  void solve();  // cpp:629-679

  // (time_step, newton_iteration, ||r||, gmres_steps or -1) of every Newton iteration: the quantities the
  // reference prints (cpp:604-606, 584) in machine-readable form
  struct Record {
    unsigned int time_step, newton_iteration;
    double residual_norm;
    int gmres_steps;
  };
  const std::vector<Record> &history() const { return history_; }
  std::vector<double> solution_owned_values() const;

protected:
  void assemble_system();         // cpp:178-378
  void solve_system();            // cpp:561-588
  void assemble_stokes_system();  // cpp:380-531
  void solve_stokes_system();     // cpp:533-559
  void solve_newton();            // cpp:590-627
  void output(const unsigned int &time_step, const double &time) const;  // cpp:681-728

  const unsigned int mpi_size, mpi_rank;
  ConditionalOStream pcout;

  // problem definition: reference constants (hpp:703-709, 438, 473-474) unless overridden by NS_* variables
  double nu = 0.001, rho = 1, p_out = 10, g = 0.0, u_m = 1.5, H = 0.41, inlet_y0 = 0.0;
  double time = 0.0;
  const double T;
  const unsigned int degree_velocity, degree_pressure;
  const double deltat;

private:
  void push_params(bool stokes);
  void dirichlet(bool stokes, std::vector<int32_t> &dofs, std::vector<double> &vals) const;
  [[noreturn]] void fail(const std::string &what) const;

  std::string mesh_path, inlet_time_mode, preconditioner, output_dir;
  int surface_entity = -1, refine_levels = 0, neumann_id = 10, inlet_id = 11;
  std::vector<int> wall_ids{12, 13}, box_tags;
  bool stokes_init = false;
  unsigned int current_step = 0;
  nst_mesh *mesh = nullptr;
  nst_dofs *dofs = nullptr;
  nst_part *part = nullptr;
  nsg_ctx *dev = nullptr;
  int64_t n_own = 0, n_own_u = 0, u_lo = 0, p_lo = 0, n_u_global = 0, n_p_global = 0;
  std::vector<Record> history_;
};

#endif
