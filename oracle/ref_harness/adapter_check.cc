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
 *
 * Built twice by ../Makefile: against the CPU oracle (oracle_shim.h renames fdal_* to fdalo_*)
 * -> _ref/libadapter_oracle.so, and against the CUDA library -> _ref/libadapter_cuda.so.
 * Return codes: 0 ok, 1 SolverControl::NoConvergence was thrown, 2 any other exception.
 */
#define FDAL_DEALII_NO_TRILINOS
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
}
