// host_finalize_hooks.cpp — C entry points over csrc/host_finalize.h for the CPU tests
// (tests/test_host_finalize.py): the same host code fdal_finalize runs, callable without a GPU.
#include "host_finalize.h"

#include <cstring>

using namespace fdal;

extern "C" {

// T = A^T (stable).  Outputs are caller-allocated: t_rp[nc + 1], t_ci[nnz], t_v[nnz].
void fdal_hostfin_transpose(int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp, const int32_t *ci, const double *v,
                            int64_t min_parallel_nnz, int32_t *t_rp, int32_t *t_ci, double *t_v) {
  HostCsr A, T;
  fill_host_csr(A, nr, nc, nnz, rp, ci, v);
  host_transpose(A, T, min_parallel_nnz);
  std::memcpy(t_rp, T.rp.data(), ((size_t)nc + 1) * sizeof(int32_t));
  if (nnz) {
    std::memcpy(t_ci, T.ci.data(), (size_t)nnz * sizeof(int32_t));
    std::memcpy(t_v, T.v.data(), (size_t)nnz * sizeof(double));
  }
}

// CSR -> BSR.  Returns the number of blocks (-1: not blocked).  Pass brp[nr / b + 1]; bcj / bv may be NULL
// (count only) or hold at least the returned number of blocks / blocks * b * b.
int64_t fdal_hostfin_bsr(int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp, const int32_t *ci, const double *v,
                         int32_t b, double max_fill, int32_t *brp, int32_t *bcj, double *bv) {
  HostCsr A;
  fill_host_csr(A, nr, nc, nnz, rp, ci, v);
  IntBuf rp_, cj_;
  DblBuf v_;
  if (!host_bsr_convert(A, b, max_fill, rp_, cj_, v_)) return -1;
  std::memcpy(brp, rp_.data(), rp_.size() * sizeof(int32_t));
  if (bcj && !cj_.empty()) std::memcpy(bcj, cj_.data(), cj_.size() * sizeof(int32_t));
  if (bv && !v_.empty()) std::memcpy(bv, v_.data(), v_.size() * sizeof(double));
  return (int64_t)cj_.size();
}

// Chebyshev plan of an exact mass solve from the (r.z, p.Ap) history of a Jacobi-PCG run.  Returns the iteration
// count (0: no plan); coef_out must hold 2 * cap doubles.
int32_t fdal_hostfin_cheb_plan(int32_t n_hist, const double *rho, const double *pv, int32_t cap, double *lo, double *hi,
                               double *coef_out) {
  CgHistory h;
  h.rho.assign(rho, rho + n_hist);
  h.pv.assign(pv, pv + n_hist);
  std::vector<double> coef;
  int its = 0;
  if (!chebyshev_plan(h, cap, lo, hi, &its, coef)) return 0;
  std::memcpy(coef_out, coef.data(), coef.size() * sizeof(double));
  return its;
}

}  // extern "C"
