/*
 * TEST INFRASTRUCTURE.  Compiles include/fdal_dealii.h (the reference-side binding of
 * INTEGRATION.md) against the stand-in deal.II types, TOGETHER WITH the reference's own
 * augmented_lagrangian_preconditioner.h (unmodified, from /root/reference), and exposes C entry
 * points that do what a patched reference application would do:
 *
 *   adapter_reference_vmult  the REFERENCE preconditioner class, constructed from the adapter's
 *                            LinearOperators (Aug_inv, C, Ct, Bt, M, invW, Mp_inv served by the
 *                            C ABI), applied to a block vector
 *   adapter_al_vmult         fdal_dealii::ALPreconditioner::vmult (fused fdal_apply_prec)
 *   adapter_solve            fdal_dealii::solve  (replaces solver_fgmres.solve(AA, x, b, P))
 *   adapter_export_csr       fdal_dealii::export_csr from a dealii::SparseMatrix
 *   adapter_to_control       fdal_dealii::to_control
 *   adapter_export_amg       fdal_dealii::export_amg from a TrilinosWrappers::PreconditionAMG whose ML
 *                            hierarchy (stand-in types of ref_harness/trilinos_stub, laid out like ML's:
 *                            Amat[l], Pmat[l+1], Rmat[l], Amat[l].lambda_max) is filled from the caller's levels
 *
 * Built twice by ../Makefile: against the CPU oracle (oracle_shim.h renames fdal_* to fdalo_*)
 * -> _ref/libadapter_oracle.so, and against the CUDA library -> _ref/libadapter_cuda.so.
 * Return codes: 0 ok, 1 SolverControl::NoConvergence was thrown, 2 any other exception.
 */
#define FDAL_STUB_TRILINOS /* export_amg compiles against ref_harness/trilinos_stub (ML / Epetra stand-ins) */
#include <augmented_lagrangian_preconditioner.h>

#include "fdal_dealii.h"

#include <cstring>

namespace {
using Vec = dealii::Vector<double>;
using BVec = dealii::BlockVector<double>;

BVec make_block(const int n_blocks, const int64_t *sizes, const double *src) {
  std::vector<std::size_t> bs(sizes, sizes + n_blocks);
  BVec v(bs);
  if (src)
    for (int b = 0; b < n_blocks; ++b) {
      std::memcpy(v.block(b).begin(), src, v.block(b).size() * sizeof(double));
      src += v.block(b).size();
    }
  return v;
}
void store(const BVec &v, double *dst) {
  for (unsigned int b = 0; b < v.n_blocks(); ++b) {
    std::memcpy(dst, v.block(b).begin(), v.block(b).size() * sizeof(double));
    dst += v.block(b).size();
  }
}
template <class F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const dealii::SolverControl::NoConvergence &) {
    return 1;
  } catch (const std::exception &) {
    return 2;
  }
}
}  // namespace

extern "C" {

int adapter_reference_vmult(fdal_ctx *ctx, int kind, double gamma, double gamma_grad_div, const int64_t *sizes,
                            const double *u_in, double *v_out) {
  return guarded([&] {
    const bool two = kind == FDAL_KIND_LAPLACE;
    const unsigned int n0 = sizes[0], n1 = sizes[1], nl = two ? sizes[1] : sizes[2];
    const BVec u = make_block(two ? 2 : 3, sizes, u_in);
    BVec v = make_block(two ? 2 : 3, sizes, nullptr);
    using namespace fdal_dealii;
    const auto Aug_inv = augmented_inverse(ctx, FDAL_AMG_A11, n0);
    const auto Ct = matrix_operator(ctx, FDAL_MAT_CT, false, n0, nl);
    const auto C = matrix_operator(ctx, FDAL_MAT_CT, true, nl, n0);
    const auto invW = winv_operator(ctx, nl);
    switch (kind) {
      case FDAL_KIND_LAPLACE: {
        const BlockPreconditionerAugmentedLagrangian P(Aug_inv, C, Ct, invW, gamma);
        P.vmult(v, u);
      } break;
      case FDAL_KIND_STOKES: {
        const BlockPreconditionerAugmentedLagrangianStokes P(Aug_inv, matrix_operator(ctx, FDAL_MAT_BT, false, n0, n1), Ct,
                                                             invW, mp_inv_operator(ctx, n1), gamma, gamma_grad_div);
        P.vmult(v, u);
      } break;
      case FDAL_KIND_STOKES_DIAG_MINRES: {
        const BlockPreconditionerAugmentedLagrangianDiagonal P(Aug_inv, invW, mp_inv_operator(ctx, n1), gamma,
                                                               gamma_grad_div);
        P.vmult(v, u);
      } break;
      case FDAL_KIND_ELLIPTIC_IDEAL: {
        const EllipticInterfacePreconditioners::BlockTriangularALPreconditioner P(
            augmented_block_inverse(ctx, n0, n1, nl), C, matrix_operator(ctx, FDAL_MAT_M, false, n1, nl), invW, gamma);
        P.vmult(v, u);
      } break;
      case FDAL_KIND_ELLIPTIC_MODIFIED: {
        const EllipticInterfacePreconditioners::BlockTriangularALPreconditionerModified P(
            C, matrix_operator(ctx, FDAL_MAT_M, false, n1, nl), invW, gamma, Aug_inv,
            augmented_inverse(ctx, FDAL_AMG_A22, n1));
        P.vmult(v, u);
      } break;
      default:
        throw std::runtime_error("unknown kind");
    }
    store(v, v_out);
  });
}

int adapter_al_vmult(fdal_ctx *ctx, int n_blocks, const int64_t *sizes, const double *u_in, double *v_out,
                     int inner_its[2]) {
  return guarded([&] {
    const BVec u = make_block(n_blocks, sizes, u_in);
    BVec v = make_block(n_blocks, sizes, nullptr);
    const fdal_dealii::ALPreconditioner P(ctx);
    P.vmult(v, u);
    inner_its[0] = P.last_inner_iterations[0];
    inner_its[1] = P.last_inner_iterations[1];
    store(v, v_out);
  });
}

int adapter_solve(fdal_ctx *ctx, int n_blocks, const int64_t *sizes, const double *rhs_in, double *x_inout,
                  fdal_solve_info *info) {
  return guarded([&] {
    const BVec rhs = make_block(n_blocks, sizes, rhs_in);
    BVec x = make_block(n_blocks, sizes, x_inout);
    fdal_dealii::solve(ctx, x, rhs, info);
    store(x, x_inout);
  });
}

int adapter_export_csr(fdal_ctx *ctx, int matrix_id, int64_t rows, int64_t cols, const int64_t *rp, const int32_t *ci,
                       const double *val) {
  return guarded([&] {
    std::vector<std::size_t> rowstart(rp, rp + rows + 1);
    std::vector<unsigned int> colnums(ci, ci + rp[rows]);
    std::vector<double> v(val, val + rp[rows]);
    const dealii::SparseMatrix<double> A(rows, cols, std::move(rowstart), std::move(colnums), std::move(v));
    fdal_dealii::export_csr(ctx, matrix_id, A);
  });
}

void adapter_to_control(int type, unsigned int max_steps, double tol, double reduce, fdal_control *out) {
  if (type == FDAL_CONTROL_REDUCTION)
    *out = fdal_dealii::to_control(dealii::ReductionControl(max_steps, tol, reduce));
  else if (type == FDAL_CONTROL_ITERATION_NUMBER)
    *out = fdal_dealii::to_control(dealii::IterationNumberControl(max_steps, tol));
  else
    *out = fdal_dealii::to_control(dealii::SolverControl(max_steps, tol));
}

/* levels: n_levels operators A_l; P[l] (n_l x n_{l+1}) and R[l] (n_{l+1} x n_l) for l < n_levels - 1, all CSR with
 * int32 row pointers, concatenated: mats = [A_0, P_0, R_0, A_1, P_1, R_1, ..., A_L] */
int adapter_export_amg(fdal_ctx *ctx, int which, int n_levels, const int32_t *n_rows, const int32_t *n_cols,
                       const int32_t *const *rp, const int32_t *const *ci, const double *const *val,
                       const double *lambda_max, int sweeps, double alpha) {
  return guarded([&] {
    std::vector<MlStubCsr> store(3 * n_levels);
    std::vector<ML_Operator> Amat(n_levels), Pmat(n_levels), Rmat(n_levels);
    auto fill = [&](int slot, ML_Operator &op) {
      MlStubCsr &m = store[slot];
      m.n_rows = n_rows[slot];
      m.n_cols = n_cols[slot];
      m.rp.assign(rp[slot], rp[slot] + m.n_rows + 1);
      m.ci.assign(ci[slot], ci[slot] + m.rp[m.n_rows]);
      m.v.assign(val[slot], val[slot] + m.rp[m.n_rows]);
      op.data = &m;
      op.invec_leng = m.n_cols;
      op.outvec_leng = m.n_rows;
    };
    for (int l = 0; l < n_levels; ++l) {
      fill(3 * l, Amat[l]);
      Amat[l].lambda_max = lambda_max[l];
      if (l + 1 < n_levels) {
        fill(3 * l + 1, Pmat[l + 1]); /* ML: Pmat[l+1] prolongates level l+1 -> l */
        fill(3 * l + 2, Rmat[l]);     /* ML: Rmat[l] restricts level l -> l+1 */
      }
    }
    ML ml;
    ml.ML_num_actual_levels = ml.ML_num_levels = n_levels;
    ml.Amat = Amat.data();
    ml.Pmat = Pmat.data();
    ml.Rmat = Rmat.data();
    const dealii::TrilinosWrappers::PreconditionAMG amg(std::make_shared<ML_Epetra::MultiLevelPreconditioner>(&ml));
    fdal_dealii::export_amg(ctx, which, amg, sweeps, alpha);
  });
}
}
