// host_finalize.h — the host-side (OpenMP) data preparation of fdal_finalize: CSR copies without a
// value-initialising pass, the stable transpose (C = Ct^T, B = Bt^T, R = P^T: SparseMatrix::Tvmult of the
// reference becomes a gather), the CSR -> BSR conversion of the dim-blocked matrices (fallback of the device
// conversion) and the plan of the Chebyshev mass solves (Lanczos bounds, iteration count, coefficients).  No CUDA in here, so
// the same code is compiled into csrc/libfdal_host.so for the CPU tests (tests/test_host_finalize.py).
#pragma once
#include <omp.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <utility>
#include <vector>

namespace fdal {

// std::vector whose resize(n) leaves the new elements uninitialised.  The big host arrays (the 1.8 G-entry
// velocity block is 21 GB) are filled right after the resize by an OpenMP loop: value-initialising them first
// costs a single-threaded, page-faulting pass over the whole buffer and puts every page on one NUMA node.
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = NoInitAlloc<U>;
  };
  NoInitAlloc() = default;
  template <class U>
  NoInitAlloc(const NoInitAlloc<U> &) {}
  template <class U, class... Args>
  void construct(U *p, Args &&...args) {
    if constexpr (sizeof...(Args) == 0)
      ::new ((void *)p) U;  // default-init: no store for int / double
    else
      ::new ((void *)p) U(std::forward<Args>(args)...);
  }
};
using IntBuf = std::vector<int, NoInitAlloc<int>>;
using DblBuf = std::vector<double, NoInitAlloc<double>>;

struct HostCsr {
  int64_t nr = 0, nc = 0, nnz = 0;
  IntBuf rp, ci;
  DblBuf v;
  bool set = false;
  // halo plan (multi-GPU): columns [0, n_owned) are owned, [n_owned, n_owned + n_halo) halo
  bool has_plan = false;
  int64_t n_owned = 0, n_halo = 0;
  std::vector<int> send_counts, recv_counts, send_idx;
  int64_t owned_cols() const { return has_plan ? n_owned : nc; }
};

inline void fill_host_csr(HostCsr &h, int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp, const int32_t *ci,
                          const double *v) {
  h = HostCsr();
  h.nr = nr;
  h.nc = nc;
  h.nnz = nnz;
  h.rp.resize((size_t)nr + 1);
  h.ci.resize((size_t)nnz);
  h.v.resize((size_t)nnz);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i <= nr; ++i) h.rp[(size_t)i] = (int)rp[i];
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) {
    h.ci[(size_t)k] = ci[k];
    h.v[(size_t)k] = v[k];
  }
  h.set = true;
}

// T = A^T by a stable counting sort: within a row of T the columns ascend, i.e. the same accumulation order
// as SparseMatrix::Tvmult's row scatter.  Parallel form (A.nnz >= min_parallel_nnz): every task owns a
// contiguous range of T's rows (cut by entry count) and scans A for them — the scan is sequential reads, the
// scattered writes (the expensive part) are split, and the order inside a row of T is the serial one.
inline void host_transpose(const HostCsr &A, HostCsr &T, int64_t min_parallel_nnz = (int64_t)1 << 22) {
  T.nr = A.nc;
  T.nc = A.nr;
  T.nnz = A.nnz;
  T.rp.assign((size_t)T.nr + 1, 0);
  T.ci.resize((size_t)A.nnz);
  T.v.resize((size_t)A.nnz);
  for (int64_t k = 0; k < A.nnz; ++k) T.rp[(size_t)A.ci[(size_t)k] + 1]++;
  for (int64_t i = 0; i < T.nr; ++i) T.rp[(size_t)i + 1] += T.rp[(size_t)i];
  const int ntasks = A.nnz < min_parallel_nnz ? 1 : std::max(1, omp_get_max_threads());
  std::vector<int64_t> cut((size_t)ntasks + 1, T.nr);
  cut[0] = 0;
  for (int t = 1; t < ntasks; ++t) {
    const int target = (int)((double)A.nnz * t / ntasks);
    const int64_t at = std::lower_bound(T.rp.begin(), T.rp.end(), target) - T.rp.begin();
    cut[(size_t)t] = std::min<int64_t>(std::max(at, cut[(size_t)t - 1]), T.nr);
  }
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < ntasks; ++t) {
    const int lo = (int)cut[(size_t)t], hi = (int)cut[(size_t)t + 1];
    if (lo >= hi) continue;
    std::vector<int> pos(T.rp.begin() + lo, T.rp.begin() + hi);
    for (int64_t i = 0; i < A.nr; ++i)
      for (int k = A.rp[(size_t)i]; k < A.rp[(size_t)i + 1]; ++k) {
        const int j = A.ci[(size_t)k];
        if (j < lo || j >= hi) continue;
        const int q = pos[(size_t)(j - lo)]++;
        T.ci[(size_t)q] = (int)i;
        T.v[(size_t)q] = A.v[(size_t)k];
      }
  }
  T.set = true;
}

// CSR -> BSR with b x b blocks stored contiguously (row-major inside a block), block columns sorted and
// merged per block row (duplicate entries are summed in entry order).  Returns false — nothing built — when
// the shape does not block, the block count overflows 32 bits or blocking would store more than
// `max_fill` x the scalar non-zeros as explicit zeros.
inline bool host_bsr_convert(const HostCsr &h, int b, double max_fill, IntBuf &brp, IntBuf &bcj, DblBuf &bv) {
  if (b < 2 || b > 3 || h.nr % b || h.nc % b || h.owned_cols() % b || h.nr == 0) return false;
  const int64_t nbr = h.nr / b;
  brp.assign((size_t)nbr + 1, 0);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 2048)
    for (int64_t I = 0; I < nbr; ++I) {
      tmp.clear();
      for (int r = 0; r < b; ++r)
        for (int k = h.rp[(size_t)(I * b + r)]; k < h.rp[(size_t)(I * b + r) + 1]; ++k) tmp.push_back(h.ci[(size_t)k] / b);
      std::sort(tmp.begin(), tmp.end());
      brp[(size_t)I + 1] = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  int64_t nblk = 0;
  for (int64_t I = 0; I < nbr; ++I) {
    nblk += brp[(size_t)I + 1];
    if (nblk >= (int64_t)std::numeric_limits<int>::max() / (b * b)) return false;
    brp[(size_t)I + 1] = (int)nblk;
  }
  if ((double)nblk * b * b > max_fill * (double)h.nnz) return false;  // too much zero fill: stay scalar
  bcj.resize((size_t)nblk);
  bv.resize((size_t)nblk * b * b);  // zeroed block row by block row inside the parallel loop (first touch)
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(dynamic, 2048)
    for (int64_t I = 0; I < nbr; ++I) {
      tmp.clear();
      for (int r = 0; r < b; ++r)
        for (int k = h.rp[(size_t)(I * b + r)]; k < h.rp[(size_t)(I * b + r) + 1]; ++k) tmp.push_back(h.ci[(size_t)k] / b);
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      const int k0 = brp[(size_t)I];
      const int nb = (int)tmp.size();
      for (int k = 0; k < nb; ++k) bcj[(size_t)k0 + k] = tmp[(size_t)k];
      double *vb = bv.data() + (size_t)k0 * b * b;
      std::fill(vb, vb + (size_t)nb * b * b, 0.0);
      for (int r = 0; r < b; ++r)
        for (int k = h.rp[(size_t)(I * b + r)]; k < h.rp[(size_t)(I * b + r) + 1]; ++k) {
          const int J = h.ci[(size_t)k] / b, q = h.ci[(size_t)k] % b;
          const int pos = (int)(std::lower_bound(tmp.begin(), tmp.end(), J) - tmp.begin());
          vb[(size_t)pos * b * b + (r * b + q)] += h.v[(size_t)k];
        }
    }
  }
  return true;
}

// ---- exact mass inverses in Chebyshev form: everything that is decided on the host ------------------------
struct CgHistory {
  std::vector<double> rho, pv;  // r.z and p.Ap of every iteration of the calibration CG
};
// extreme eigenvalues of the symmetric tridiagonal matrix (d, e) by Sturm bisection
inline void tridiag_extremes(const std::vector<double> &d, const std::vector<double> &e, double *lo_out, double *hi_out) {
  const int n = (int)d.size();
  double gl = d[0], gu = d[0];
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(e[(size_t)i - 1]) : 0.0) + (i + 1 < n ? std::fabs(e[(size_t)i]) : 0.0);
    gl = std::min(gl, d[(size_t)i] - r);
    gu = std::max(gu, d[(size_t)i] + r);
  }
  auto count_below = [&](double x) {  // eigenvalues < x
    int cnt = 0;
    double q = d[0] - x;
    for (int i = 0;; ++i) {
      if (q < 0.0) ++cnt;
      if (i + 1 == n) break;
      if (q == 0.0) q = 1e-300;
      q = d[(size_t)i + 1] - x - e[(size_t)i] * e[(size_t)i] / q;
    }
    return cnt;
  };
  auto kth = [&](int k) {  // smallest x with count_below(x) >= k, i.e. the k-th eigenvalue (1-based)
    double a = gl, b = gu;
    for (int i = 0; i < 200 && b - a > 1e-15 * std::max(std::fabs(a), std::fabs(b)); ++i) {
      const double mid = 0.5 * (a + b);
      if (count_below(mid) >= k)
        b = mid;
      else
        a = mid;
    }
    return 0.5 * (a + b);
  };
  *lo_out = kth(1);
  *hi_out = kth(n);
}
// Ritz values of D^-1 M from the CG coefficients of the calibration solve (Lanczos connection:
// T_jj = 1/alpha_j + beta_j/alpha_{j-1}, T_{j,j+1} = sqrt(beta_{j+1})/alpha_j)
inline bool lanczos_bounds(const CgHistory &h, double *lo, double *hi) {
  std::vector<double> d, e;
  double alpha_prev = 0.0;
  for (size_t j = 0; j < h.rho.size() && j < h.pv.size(); ++j) {
    const double rho = h.rho[j], pv = h.pv[j];
    if (!(rho > 0.0) || !(pv > 0.0) || !std::isfinite(rho) || !std::isfinite(pv)) break;
    const double alpha = rho / pv;
    const double beta = j > 0 ? rho / h.rho[j - 1] : 0.0;
    if (j > 0) e.push_back(std::sqrt(beta) / alpha_prev);
    d.push_back(1.0 / alpha + (j > 0 ? beta / alpha_prev : 0.0));
    alpha_prev = alpha;
  }
  if (d.size() < 3) return false;
  e.resize(d.size() - 1);
  tridiag_extremes(d, e, lo, hi);
  return *lo > 0.0 && *hi > *lo && std::isfinite(*hi);
}
// The fixed-count Chebyshev iteration on D^-1 M that replaces the exact mass solve: spectral interval = the Ritz
// interval widened by 3 % on either side (Ritz values lie inside the spectrum; costs ~2 iterations), iteration
// count from the Chebyshev error bound 2 q^k / (1 + q^2k) <= 1e-16 (+2), coefficients of
//   r = b - M x ; d = c1_k d + c2_k D^-1 r ; x += d        (coef[2k] = c1_k, coef[2k+1] = c2_k, c1_0 = 0).
// Returns false (keep the Jacobi-PCG) when the bounds are unusable or the count exceeds `cap`.
inline bool chebyshev_plan(const CgHistory &h, int cap, double *lo_out, double *hi_out, int *its_out,
                           std::vector<double> &coef) {
  double lo, hi;
  if (!lanczos_bounds(h, &lo, &hi)) return false;
  lo *= 0.97;
  hi *= 1.03;
  const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), s1 = theta / delta;
  const double sk = std::sqrt(hi / lo), q = (sk - 1.0) / (sk + 1.0);
  int its = (int)std::ceil(std::log(2.0e16) / -std::log(q)) + 2;
  if (its > cap) return false;
  its = std::max(its, 2);
  coef.assign(2 * (size_t)its, 0.0);
  coef[1] = 1.0 / theta;
  double rho = 1.0 / s1;
  for (int k = 1; k < its; ++k) {
    const double rho1 = 1.0 / (2.0 * s1 - rho);
    coef[2 * (size_t)k] = rho1 * rho;
    coef[2 * (size_t)k + 1] = 2.0 * rho1 / delta;
    rho = rho1;
  }
  *lo_out = lo;
  *hi_out = hi;
  *its_out = its;
  return true;
}

}  // namespace fdal
