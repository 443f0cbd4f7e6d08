/*
 * fdal_dealii.h — header-only adapter between deal.II / Trilinos objects and the C ABI of
 * fdal.h: the reference-side binding a maintainer adds next to
 * augmented_lagrangian_preconditioner.h.  deal.II (>= 9.6), Trilinos ML and UMFPACK are not
 * installed in the build image, so it has never been compiled against the real
 * libraries.  What IS compiled and run here (tests/test_dealii_adapter.py): all of it, against the
 * stand-in deal.II types of oracle/ref_harness/dealii_stub and — for export_amg — the ML / Epetra
 * stand-ins of oracle/ref_harness/trilinos_stub (real type and member names, minimal behaviour),
 * together with the reference's own preconditioner classes: the reference class built from this
 * adapter's LinearOperators reproduces fdal_apply_prec, and a context filled by export_amg applies the
 * same V-cycle as one filled through fdal_amg_set_level.  -DFDAL_DEALII_NO_TRILINOS drops export_amg.
 *
 * What it provides:
 *   fdal_dealii::export_csr          dealii::SparseMatrix<double>            -> fdal_set_csr
 *   fdal_dealii::export_amg          TrilinosWrappers::PreconditionAMG (ML)  -> fdal_amg_set_level/_coarse
 *   fdal_dealii::to_control          dealii::SolverControl family            -> fdal_control
 *   fdal_dealii::ALPreconditioner    duck-typed `vmult(BlockVector&, const BlockVector&) const`
 *                                    usable as P in solver_fgmres.solve(AA, x, b, P)
 *                                    (augmented_lagrangian_preconditioner.h:28,62,95,130,186)
 *   fdal_dealii::augmented_operator / augmented_inverse   LinearOperator<Vector<double>> for
 *                                    Aug and Aug_inv (immersed_laplace.cc:884, 911-912)
 *   fdal_dealii::solve               replaces solver_fgmres.solve(AA, x, b, P)
 *                                    (immersed_laplace.cc:943, stokes_immersed_boundary.cc:1073,
 *                                     elliptic_interface.cc:905/947); throws
 *                                    SolverControl::NoConvergence exactly where deal.II would.
 */
#ifndef FDAL_DEALII_H
#define FDAL_DEALII_H

#include <deal.II/base/exceptions.h>
#include <deal.II/lac/block_vector.h>
#include <deal.II/lac/linear_operator.h>
#include <deal.II/lac/solver_control.h>
#include <deal.II/lac/sparse_matrix.h>
#include <deal.II/lac/trilinos_precondition.h>
#include <deal.II/lac/vector.h>

#ifndef FDAL_DEALII_NO_TRILINOS
#include <ml_MultiLevelPreconditioner.h>
#include <ml_epetra_utils.h>
#endif

#include <memory>
#include <vector>

#include "fdal.h"

namespace fdal_dealii {

inline void check(fdal_ctx *ctx, const int status) {
  AssertThrow(status == FDAL_OK, dealii::ExcMessage(fdal_last_error(ctx)));
}

/* deal.II keeps std::size_t rowstart / unsigned int colnums / double val, with the diagonal first
 * in each row of a square matrix; the library keeps the order it is given. */
inline void export_csr(fdal_ctx *ctx, const int matrix_id, const dealii::SparseMatrix<double> &A) {
  std::vector<int64_t> rp(A.m() + 1);
  std::vector<int32_t> ci;
  std::vector<double> v;
  ci.reserve(A.n_nonzero_elements());
  v.reserve(A.n_nonzero_elements());
  for (unsigned int r = 0; r < A.m(); ++r) {
    rp[r] = static_cast<int64_t>(ci.size());
    for (auto it = A.begin(r); it != A.end(r); ++it) {
      ci.push_back(static_cast<int32_t>(it->column()));
      v.push_back(it->value());
    }
  }
  rp[A.m()] = static_cast<int64_t>(ci.size());
  check(ctx, fdal_set_csr(ctx, matrix_id, A.m(), A.n(), static_cast<int64_t>(ci.size()), rp.data(), ci.data(),
                          v.data()));
}

#ifndef FDAL_DEALII_NO_TRILINOS
struct OwnedCsr {
  std::vector<int64_t> rp;
  std::vector<int32_t> ci;
  std::vector<double> v;
  fdal_csr_view view{};
};
inline void from_epetra(const Epetra_CrsMatrix &E, OwnedCsr &out) {
  const int n = E.NumMyRows();
  out.rp.assign(n + 1, 0);
  out.ci.clear();
  out.v.clear();
  for (int r = 0; r < n; ++r) {
    int cnt;
    double *vals;
    int *idx;
    E.ExtractMyRowView(r, cnt, vals, idx);
    for (int k = 0; k < cnt; ++k) {
      out.ci.push_back(E.GCID(idx[k]));
      out.v.push_back(vals[k]);
    }
    out.rp[r + 1] = static_cast<int64_t>(out.ci.size());
  }
  out.view = {n, E.NumGlobalCols(), static_cast<int64_t>(out.ci.size()), out.rp.data(), out.ci.data(), out.v.data()};
}

/* ML hierarchy -> exchange format: per level A_l, P_l (level l+1 -> l), R_l, the
 * Chebyshev eigenvalue estimate ML computed, degree = smoother sweeps, ratio = "smoother: Chebyshev
 * alpha" (deal.II sets 10).  ML numbers levels fine -> coarse as 0..L-1 for "MGV" with
 * "increasing or decreasing" = "increasing". */
inline void export_amg(fdal_ctx *ctx, const int which, const dealii::TrilinosWrappers::PreconditionAMG &amg,
                       const int smoother_sweeps = 2, const double chebyshev_alpha = 10.0) {
  const auto *mlp = dynamic_cast<const ML_Epetra::MultiLevelPreconditioner *>(&amg.trilinos_operator());
  AssertThrow(mlp != nullptr, dealii::ExcMessage("PreconditionAMG does not hold an ML preconditioner"));
  const ML *ml = mlp->GetML();
  const int n_levels = ml->ML_num_actual_levels;
  for (int l = 0; l < n_levels; ++l) {
    Epetra_CrsMatrix *A = nullptr, *P = nullptr, *R = nullptr;
    int max_nz;
    double cpu;
    ML_Operator2EpetraCrsMatrix(&ml->Amat[l], A, max_nz, false, cpu);
    OwnedCsr a, p, r;
    from_epetra(*A, a);
    if (l + 1 < n_levels) {
      ML_Operator2EpetraCrsMatrix(&ml->Pmat[l + 1], P, max_nz, false, cpu);  /* maps level l+1 -> l */
      ML_Operator2EpetraCrsMatrix(&ml->Rmat[l], R, max_nz, false, cpu);      /* maps level l -> l+1 */
      from_epetra(*P, p);
      from_epetra(*R, r);
      const double lambda_max = ml->Amat[l].lambda_max;
      check(ctx, fdal_amg_set_level(ctx, which, l, &a.view, &p.view, &r.view, /*inv_diag=*/nullptr, lambda_max,
                                    smoother_sweeps, chebyshev_alpha));
    } else {
      check(ctx, fdal_amg_set_coarse(ctx, which, l, &a.view)); /* Amesos-KLU level: direct solve */
    }
    delete A;
    delete P;
    delete R;
  }
}

#endif /* FDAL_DEALII_NO_TRILINOS */

inline fdal_control to_control(const dealii::SolverControl &c) {
  fdal_control out{FDAL_CONTROL_SOLVER, static_cast<int32_t>(c.max_steps()), c.tolerance(), 0.0};
  if (const auto *r = dynamic_cast<const dealii::ReductionControl *>(&c)) {
    out.type = FDAL_CONTROL_REDUCTION;
    out.reduce = r->reduction();
  } else if (dynamic_cast<const dealii::IterationNumberControl *>(&c) != nullptr) {
    out.type = FDAL_CONTROL_ITERATION_NUMBER;
  }
  return out;
}

namespace internal {
inline void gather(const dealii::BlockVector<double> &src, std::vector<double> &flat) {
  flat.resize(src.size());
  std::size_t o = 0;
  for (unsigned int b = 0; b < src.n_blocks(); ++b) {
    std::copy(src.block(b).begin(), src.block(b).end(), flat.begin() + o);
    o += src.block(b).size();
  }
}
inline void scatter(const std::vector<double> &flat, dealii::BlockVector<double> &dst) {
  std::size_t o = 0;
  for (unsigned int b = 0; b < dst.n_blocks(); ++b) {
    std::copy(flat.begin() + o, flat.begin() + o + dst.block(b).size(), dst.block(b).begin());
    o += dst.block(b).size();
  }
}
}  // namespace internal

/* Any of the five AL preconditioners (the context's kind selects which): a valid P for
 * solver.solve(AA, x, b, P).  One host<->device round trip per vmult ("parity mode"). */
class ALPreconditioner {
 public:
  explicit ALPreconditioner(fdal_ctx *ctx_) : ctx(ctx_) {}
  void vmult(dealii::BlockVector<double> &dst, const dealii::BlockVector<double> &src) const {
    internal::gather(src, u);
    v.resize(u.size());
    int its[2] = {0, 0};
    const int st = fdal_apply_prec(ctx, u.data(), v.data(), its);
    if (st == FDAL_ERR_INNER_NO_CONVERGENCE || st == FDAL_ERR_MASS_NO_CONVERGENCE)
      throw dealii::SolverControl::NoConvergence(its[0], 0.); /* what the inner SolverCG throws */
    check(ctx, st);
    internal::scatter(v, dst);
    last_inner_iterations[0] = its[0];
    last_inner_iterations[1] = its[1];
  }
  mutable int last_inner_iterations[2] = {0, 0};

 private:
  fdal_ctx *ctx;
  mutable std::vector<double> u, v;
};

/* Aug = K + gamma * Ct * invW * C as a LinearOperator (immersed_laplace.cc:884) */
inline dealii::LinearOperator<dealii::Vector<double>> augmented_operator(fdal_ctx *ctx, const int which,
                                                                         const unsigned int n) {
  dealii::LinearOperator<dealii::Vector<double>> op;
  op.vmult = [ctx, which](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    check(ctx, fdal_apply_aug(ctx, which, x.begin(), y.begin()));
  };
  op.vmult_add = [ctx, which](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    dealii::Vector<double> t(y.size());
    check(ctx, fdal_apply_aug(ctx, which, x.begin(), t.begin()));
    y += t;
  };
  op.Tvmult = op.vmult; /* symmetric */
  op.Tvmult_add = op.vmult_add;
  op.reinit_range_vector = op.reinit_domain_vector = [n](dealii::Vector<double> &v, bool fast) { v.reinit(n, fast); };
  return op;
}
/* Aug_inv = inverse_operator(Aug, SolverCG, AMG) (immersed_laplace.cc:911-912): zero initial guess,
 * the configured inner control, NoConvergence on failure */
inline dealii::LinearOperator<dealii::Vector<double>> augmented_inverse(fdal_ctx *ctx, const int which,
                                                                        const unsigned int n) {
  dealii::LinearOperator<dealii::Vector<double>> op;
  op.vmult = [ctx, which](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    int its = 0;
    const int st = fdal_apply_aug_inv(ctx, which, x.begin(), y.begin(), &its);
    if (st == FDAL_ERR_INNER_NO_CONVERGENCE) throw dealii::SolverControl::NoConvergence(its, 0.);
    check(ctx, st);
  };
  op.reinit_range_vector = op.reinit_domain_vector = [n](dealii::Vector<double> &v, bool fast) { v.reinit(n, fast); };
  return op;
}

/* The "ideal" elliptic preconditioner takes Aug_inv as a LinearOperator on the first TWO blocks
 * (elliptic_interface.cc:930-942; augmented_lagrangian_preconditioner.h:118,152).  With a zero
 * multiplier residual the preconditioner's own vmult is exactly that block solve (header lines
 * 135-155 with u.block(2) = 0), so it is served by fdal_apply_prec on [x0 | x1 | 0]. */
inline dealii::LinearOperator<dealii::BlockVector<double>> augmented_block_inverse(fdal_ctx *ctx, const unsigned int n0,
                                                                                   const unsigned int n1,
                                                                                   const unsigned int n_lambda) {
  dealii::LinearOperator<dealii::BlockVector<double>> op;
  auto reinit = [n0, n1](dealii::BlockVector<double> &v, bool fast) {
    v.reinit(2);
    v.block(0).reinit(n0, fast);
    v.block(1).reinit(n1, fast);
    v.collect_sizes();
  };
  op.reinit_range_vector = op.reinit_domain_vector = reinit;
  op.vmult = [ctx, n0, n1, n_lambda](dealii::BlockVector<double> &y, const dealii::BlockVector<double> &x) {
    std::vector<double> u(n0 + n1 + n_lambda, 0.0), v(n0 + n1 + n_lambda);
    std::copy(x.block(0).begin(), x.block(0).end(), u.begin());
    std::copy(x.block(1).begin(), x.block(1).end(), u.begin() + n0);
    int its[2] = {0, 0};
    const int st = fdal_apply_prec(ctx, u.data(), v.data(), its);
    if (st == FDAL_ERR_INNER_NO_CONVERGENCE) throw dealii::SolverControl::NoConvergence(its[0], 0.);
    check(ctx, st);
    std::copy(v.begin(), v.begin() + n0, y.block(0).begin());
    std::copy(v.begin() + n0, v.begin() + n0 + n1, y.block(1).begin());
  };
  return op;
}

/* linear_operator(matrix) / transpose_operator(...) of an exported block, applied on the device:
 * C, Ct, Bt, M of immersed_laplace.cc:638-642, stokes_immersed_boundary.cc:923-929,
 * elliptic_interface.cc:680-687.  `transpose` selects SparseMatrix::Tvmult. */
inline dealii::LinearOperator<dealii::Vector<double>> matrix_operator(fdal_ctx *ctx, const int matrix_id,
                                                                      const bool transpose, const unsigned int n_range,
                                                                      const unsigned int n_domain) {
  dealii::LinearOperator<dealii::Vector<double>> op;
  op.vmult = [ctx, matrix_id, transpose](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    check(ctx, fdal_spmv(ctx, matrix_id, transpose ? 1 : 0, x.begin(), y.begin()));
  };
  op.Tvmult = [ctx, matrix_id, transpose](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    check(ctx, fdal_spmv(ctx, matrix_id, transpose ? 0 : 1, x.begin(), y.begin()));
  };
  op.vmult_add = [vm = op.vmult, n_range](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    dealii::Vector<double> t(n_range);
    vm(t, x);
    y += t;
  };
  op.Tvmult_add = [tvm = op.Tvmult, n_domain](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    dealii::Vector<double> t(n_domain);
    tvm(t, x);
    y += t;
  };
  op.reinit_range_vector = [n_range](dealii::Vector<double> &v, bool fast) { v.reinit(n_range, fast); };
  op.reinit_domain_vector = [n_domain](dealii::Vector<double> &v, bool fast) { v.reinit(n_domain, fast); };
  return op;
}
/* invW (immersed_laplace.cc:849-878, stokes_immersed_boundary.cc:966-985, elliptic_interface.cc:693-739) */
inline dealii::LinearOperator<dealii::Vector<double>> winv_operator(fdal_ctx *ctx, const unsigned int m) {
  dealii::LinearOperator<dealii::Vector<double>> op;
  op.vmult = [ctx](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    check(ctx, fdal_apply_winv(ctx, x.begin(), y.begin()));
  };
  op.Tvmult = op.vmult; /* symmetric */
  op.vmult_add = op.Tvmult_add = [ctx, m](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    dealii::Vector<double> t(m);
    check(ctx, fdal_apply_winv(ctx, x.begin(), t.begin()));
    y += t;
  };
  op.reinit_range_vector = op.reinit_domain_vector = [m](dealii::Vector<double> &v, bool fast) { v.reinit(m, fast); };
  return op;
}
/* Mp_inv (stokes_immersed_boundary.cc:931-963): inner CG failure -> SolverControl::NoConvergence */
inline dealii::LinearOperator<dealii::Vector<double>> mp_inv_operator(fdal_ctx *ctx, const unsigned int n_p) {
  dealii::LinearOperator<dealii::Vector<double>> op;
  op.vmult = [ctx](dealii::Vector<double> &y, const dealii::Vector<double> &x) {
    int its = 0;
    const int st = fdal_apply_mp_inv(ctx, x.begin(), y.begin(), &its);
    if (st == FDAL_ERR_MASS_NO_CONVERGENCE) throw dealii::SolverControl::NoConvergence(its, 0.);
    check(ctx, st);
  };
  op.Tvmult = op.vmult;
  op.reinit_range_vector = op.reinit_domain_vector = [n_p](dealii::Vector<double> &v, bool fast) { v.reinit(n_p, fast); };
  return op;
}

/* Drop-in for  solver_fgmres.solve(AA, solution_block, system_rhs_block, P)  — the whole outer
 * solve stays on the device.  `control` receives last_step()/last_value() like deal.II's. */
inline void solve(fdal_ctx *ctx, dealii::BlockVector<double> &x, const dealii::BlockVector<double> &rhs,
                  fdal_solve_info *info_out = nullptr) {
  std::vector<double> b, sol;
  internal::gather(rhs, b);
  internal::gather(x, sol);
  fdal_solve_info info;
  const int st = fdal_solve(ctx, b.data(), sol.data(), &info);
  if (info_out) *info_out = info;
  if (st == FDAL_ERR_OUTER_NO_CONVERGENCE || st == FDAL_ERR_INNER_NO_CONVERGENCE || st == FDAL_ERR_MASS_NO_CONVERGENCE) {
    internal::scatter(sol, x); /* deal.II leaves the last iterate in x when it throws */
    throw dealii::SolverControl::NoConvergence(info.outer_iterations, info.final_residual);
  }
  check(ctx, st);
  internal::scatter(sol, x);
}

}  // namespace fdal_dealii
#endif /* FDAL_DEALII_H */
