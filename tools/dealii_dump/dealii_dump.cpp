// dealii_dump — the one way "bit-exact against deal.II + Trilinos" can be certified (pin 4).
//
// NOT built in this repository's image (deal.II, Trilinos and MPI are absent: DESIGN.md §6).  On a machine
// with deal.II >= 9.3.1 configured with Trilinos and MPI (e.g. the PoliMi `module load gcc-glibc dealii`
// environment the reference's README names), build it against the UNMODIFIED reference sources:
//
//     cmake -S tools/dealii_dump -B build/dealii_dump -DREFERENCE_DIR=/path/to/Navier-Stokes-dealii && cmake --build build/dealii_dump
//     cd /path/to/Navier-Stokes-dealii/build && mpirun -np 1 /path/to/build/dealii_dump/dealii_dump dump_dir
//
// The program derives from the reference's NavierStokesSolver (every member it needs is `protected`,
// src/NavierStokesSolver.hpp:664-794), runs setup() and then, for the first Newton iteration of the first
// time step, assemble_system() and solve_system() exactly as solve_newton() does (src/NavierStokesSolver.cpp:
// 590-627), writing after each stage the objects this repository restates:
//
//     cells.txt      one line per active cell in iteration order: 3 x (x y) vertex coordinates, 15 global dof ids
//     pattern.txt    CSR of the 2x2 block Jacobian in GLOBAL dof numbering: "row: col col ..." (Trilinos column order)
//     pm_pattern.txt CSR of pressure_mass (block (1,1) rows only carry entries)
//     jacobian.txt   "row col value" of every stored entry after assemble_system() (%.17g)
//     pm.txt         same for pressure_mass
//     residual.txt   residual_vector after assemble_system()
//     delta.txt      delta_owned after solve_system(); gmres.txt: last step and residual of SolverControl is not
//                    reachable from outside solve_system(), so only the increment is dumped
//
// tools/dealii_dump/compare.py loads these files and checks this repository's topology library (bit-exact
// dof maps and pattern), CPU oracle and CUDA path (1e-12 entries/residual, 1e-8 increment) against them.
#include <deal.II/base/conditional_ostream.h>

#include <filesystem>
#include <fstream>
#include <iomanip>

#include "NavierStokesSolver.hpp"

using namespace dealii;

class DumpingSolver : public NavierStokesSolver
{
public:
  using NavierStokesSolver::NavierStokesSolver;

  void
  run(const std::string &dir)
  {
    std::filesystem::create_directories(dir);
    setup();

    const unsigned int rank = Utilities::MPI::this_mpi_process(MPI_COMM_WORLD);
    const std::string  sfx  = Utilities::MPI::n_mpi_processes(MPI_COMM_WORLD) > 1 ? "." + std::to_string(rank) : "";

    {
      std::ofstream                        f(dir + "/cells.txt" + sfx);
      std::vector<types::global_dof_index> ids(fe->dofs_per_cell);
      f << std::setprecision(17);
      for (const auto &cell : dof_handler.active_cell_iterators())
        {
          if (!cell->is_locally_owned())
            continue;
          cell->get_dof_indices(ids);
          for (unsigned int v = 0; v < 3; ++v)
            f << cell->vertex(v)[0] << ' ' << cell->vertex(v)[1] << ' ';
          for (const auto id : ids)
            f << id << ' ';
          f << '\n';
        }
    }

    // what solve() does before the first solve_newton(): initial condition, first time level (cpp:646-666)
    VectorTools::interpolate(dof_handler, u_0, solution_owned);
    solution = solution_owned;
    time += deltat;
    solution_old = solution;

    assemble_system();
    dump_block_matrix(jacobian_matrix, dir + "/pattern.txt" + sfx, dir + "/jacobian.txt" + sfx);
    dump_block_matrix(pressure_mass, dir + "/pm_pattern.txt" + sfx, dir + "/pm.txt" + sfx);
    dump_block_vector(residual_vector, dir + "/residual.txt" + sfx);

    solve_system();
    dump_block_vector(delta_owned, dir + "/delta.txt" + sfx);
  }

private:
  // global index of (block b, index i inside the block): blocks are numbered one after the other (cpp:73-91)
  static types::global_dof_index
  offset_of(const TrilinosWrappers::BlockSparseMatrix &M, const unsigned int b, const bool column)
  {
    types::global_dof_index o = 0;
    for (unsigned int k = 0; k < b; ++k)
      o += column ? M.block(0, k).n() : M.block(k, 0).m();
    return o;
  }

  void
  dump_block_matrix(const TrilinosWrappers::BlockSparseMatrix &M, const std::string &pattern_file, const std::string &value_file) const
  {
    std::ofstream fp(pattern_file), fv(value_file);
    fv << std::setprecision(17);
    for (unsigned int br = 0; br < M.n_block_rows(); ++br)
      {
        const auto &first = M.block(br, 0);
        const auto  range = first.local_range();
        for (types::global_dof_index r = range.first; r < range.second; ++r)
          {
            fp << offset_of(M, br, false) + r << ':';
            for (unsigned int bc = 0; bc < M.n_block_cols(); ++bc)
              {
                const auto &B = M.block(br, bc);
                for (auto it = B.begin(r); it != B.end(r); ++it) // Epetra's stored (local-index) column order
                  {
                    fp << ' ' << offset_of(M, bc, true) + it->column();
                    fv << offset_of(M, br, false) + r << ' ' << offset_of(M, bc, true) + it->column() << ' ' << it->value() << '\n';
                  }
              }
            fp << '\n';
          }
      }
  }

  void
  dump_block_vector(const TrilinosWrappers::MPI::BlockVector &v, const std::string &file) const
  {
    std::ofstream f(file);
    f << std::setprecision(17);
    types::global_dof_index off = 0;
    for (unsigned int b = 0; b < v.n_blocks(); ++b)
      {
        const auto &blk = v.block(b);
        for (const auto i : blk.locally_owned_elements())
          f << off + i << ' ' << blk[i] << '\n';
        off += blk.size();
      }
  }
};

int
main(int argc, char *argv[])
{
  Utilities::MPI::MPI_InitFinalize mpi_init(argc, argv);
  // the constructor arguments of src/main.cpp:9-15
  const unsigned int degree_velocity = 2;
  const unsigned int degree_pressure = 1;
  const double       T               = 1.0;
  const double       deltat          = 0.05;
  DumpingSolver      problem(degree_velocity, degree_pressure, T, deltat);
  problem.run(argc > 1 ? argv[1] : "dealii_dump_out");
  return 0;
}
