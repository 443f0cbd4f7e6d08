/*
 * What a maintainer adds to immersed_laplace.cc to run the `Solver = augmented` branch on the GPU:
 * one function that takes the objects DistributedLagrangeProblem::solve() has built by line 880
 * (stiffness matrix, coupling matrix Ct, immersed mass matrix, the diagonal W^-1, the Trilinos AMG of
 * the explicit augmented block, the two solver controls) and replaces lines 880-944 — the operator
 * definitions, the preconditioner and `solver_fgmres.solve(AA, x, b, P)` — through
 * include/fdal_dealii.h.  Everything before (assembly, amg_prec.initialize) and after
 * (constraints.distribute, output) stays as it is.
 *
 * Compiled here against the stand-in deal.II / ML types (tests/test_abi.py::test_dealii_example_compiles:
 * g++ -fsyntax-only -Ioracle/ref_harness/dealii_stub -Ioracle/ref_harness/trilinos_stub -Iinclude
 * -DFDAL_STUB_TRILINOS); against a real deal.II (>= 9.6 with Trilinos) drop the two stub include paths.
 */
#include <deal.II/lac/block_vector.h>
#include <deal.II/lac/solver_control.h>
#include <deal.II/lac/sparse_matrix.h>
#include <deal.II/lac/trilinos_precondition.h>
#include <deal.II/lac/vector.h>

#include "fdal_dealii.h"

/* returns SolverControl::last_step() of the outer solve (results_data.outer_iterations, immersed_laplace.cc:956) */
inline unsigned int solve_augmented_on_gpu(const dealii::SparseMatrix<double> &stiffness_matrix,
                                           const dealii::SparseMatrix<double> &coupling_matrix, /* Ct, n x m */
                                           const dealii::SparseMatrix<double> &mass_matrix,     /* immersed M  */
                                           const dealii::Vector<double> &inverse_squares,       /* 1 / M_ii^2  */
                                           const bool use_diagonal_inverse,
                                           const dealii::TrilinosWrappers::PreconditionAMG &amg_prec,
                                           const double gamma, const dealii::SolverControl &schur_solver_control,
                                           const dealii::SolverControl &control_lagrangian,
                                           dealii::BlockVector<double> &solution_block,
                                           dealii::BlockVector<double> &system_rhs_block, const int device = 0) {
  fdal_config cfg{};
  cfg.kind = FDAL_KIND_LAPLACE;
  cfg.restart = 30; /* SolverFGMRES default max_basis_size */
  cfg.gamma = gamma;
  cfg.winv_mode = use_diagonal_inverse ? FDAL_WINV_DIAG : FDAL_WINV_EXACT_M_SQUARED; /* :869-876 */
  cfg.inner_prec = FDAL_PREC_AMG;
  cfg.device = device;
  cfg.use_graphs = 1;
  cfg.outer = fdal_dealii::to_control(schur_solver_control); /* ReductionControl(1000, 1e-10, 1e-12) */
  cfg.inner = fdal_dealii::to_control(control_lagrangian);   /* SolverControl(100, 1e-2), :907 */
  fdal_ctx *ctx = nullptr;
  AssertThrow(fdal_create(&ctx, &cfg) == FDAL_OK, dealii::ExcMessage("fdal_create failed (no CUDA device?)"));
  struct Guard {
    fdal_ctx *c;
    ~Guard() { fdal_destroy(c); }
  } guard{ctx};

  fdal_dealii::export_csr(ctx, FDAL_MAT_A, stiffness_matrix);
  fdal_dealii::export_csr(ctx, FDAL_MAT_CT, coupling_matrix);
  fdal_dealii::export_csr(ctx, FDAL_MAT_M, mass_matrix);
  if (use_diagonal_inverse)
    fdal_dealii::check(ctx, fdal_set_diag(ctx, FDAL_DIAG_W_INV, inverse_squares.size(), inverse_squares.begin()));
  fdal_dealii::export_amg(ctx, FDAL_AMG_A11, amg_prec, /*smoother_sweeps=*/2, /*Chebyshev alpha=*/10.0);
  fdal_dealii::check(ctx, fdal_finalize(ctx));

  /* rhs0 = f + gamma Ct W^-1 g (:899-905): in place on the flat block vector, then the whole outer solve */
  std::vector<double> rhs;
  fdal_dealii::internal::gather(system_rhs_block, rhs);
  fdal_dealii::check(ctx, fdal_augment_rhs(ctx, rhs.data()));
  fdal_dealii::internal::scatter(rhs, system_rhs_block);
  fdal_solve_info info;
  fdal_dealii::solve(ctx, solution_block, system_rhs_block, &info); /* throws NoConvergence like deal.II */
  return static_cast<unsigned int>(info.outer_iterations);
}
