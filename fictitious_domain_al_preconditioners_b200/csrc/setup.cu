// setup.cu — device-side pieces of the setup phase: work the reference does on the host
// once per solve, BEFORE the hot path.  Stand-alone C-ABI entry point (no fdal_ctx): host CSR in, host CSR
// out, arithmetic on the device.
//
//   fdal_assemble_al_term   the operator-form AL term  gamma * sum_q phi_i(x_q) phi_j(x_q) JxW_q  scattered into
//                           the stiffness matrix (immersed_laplace.cc:659-702, utilities.h:755-837,
//                           nitsche_bcs.cc:517-572): one thread per (quadrature point, test function), FP64
//                           atomicAdd into the CSR values (atomics-bound, as SURVEY 8(f) N3 expects)
//
// Not here: the Galerkin / augmented-block SpGEMMs of the AMG setup (N2) stay on the host
// (OpenMP Gustavson, csrc/host_setup.c); a first device design (warp per row, sequential over A's entries for
// a deterministic sum) was estimated at > 100 s for the 1.8 G-entry fine level against 9 s on 16 cores and
// was not built.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/fdal.h"

namespace {

#define CUS(call)                           \
  do {                                      \
    cudaError_t e_ = (call);                \
    if (e_ != cudaSuccess) {                \
      fprintf(stderr, "fdal setup: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return FDAL_ERR_CUDA;                 \
    }                                       \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t n) { return cudaMalloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T)); }
  cudaError_t upload(const T *h, size_t n) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess) return e;
    return n ? cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice) : cudaSuccess;
  }
};

// ---------------------------------------------------------------------------------------------- N3
// thread t = (point q, local test function i): row = dofs[q][i]; for every local trial function j the entry
// (row, dofs[q][j]) of the CSR matrix receives weight[q] * phi[q][i] * phi[q][j].  Rows are searched linearly
// (deal.II rows are "diagonal first, then ascending": not sorted; FE rows are short).  A dof index < 0 marks a
// constrained / absent local dof and is skipped.  missing[0] counts entries the sparsity pattern lacks.
__global__ void k_al_term_scatter(long long n_points, int dpc, const int *__restrict__ dofs,
                                  const double *__restrict__ phi, const double *__restrict__ weight,
                                  const long long *__restrict__ rp, const int *__restrict__ ci, double *val,
                                  unsigned long long *missing) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_points * dpc) return;
  const long long q = t / dpc;
  const int i = (int)(t - q * dpc);
  const int row = dofs[q * dpc + i];
  if (row < 0) return;
  const double wi = weight[q] * phi[q * dpc + i];
  if (wi == 0.0) return;
  const long long k0 = rp[row], k1 = rp[row + 1];
  for (int j = 0; j < dpc; ++j) {
    const int col = dofs[q * dpc + j];
    if (col < 0) continue;
    const double a = wi * phi[q * dpc + j];
    if (a == 0.0) continue;
    long long k = k0;
    while (k < k1 && ci[k] != col) ++k;
    if (k < k1)
      atomicAdd(val + k, a);
    else
      atomicAdd(missing, 1ull);
  }
}

}  // namespace

extern "C" {

int fdal_assemble_al_term(int device, int64_t n_rows, const int64_t *row_ptr, const int32_t *col, double *val_inout,
                          int64_t n_points, int32_t dofs_per_cell, const int32_t *point_dofs, const double *point_phi,
                          const double *weight, int64_t *n_missing_out) {
  if (n_rows < 0 || !row_ptr || n_points < 0 || dofs_per_cell < 1 || dofs_per_cell > 512) return FDAL_ERR_INVALID;
  const int64_t nnz = row_ptr[n_rows];
  if (nnz && (!col || !val_inout)) return FDAL_ERR_INVALID;
  if (n_points && (!point_dofs || !point_phi || !weight)) return FDAL_ERR_INVALID;
  for (int64_t k = 0; k < n_points * dofs_per_cell; ++k)
    if (point_dofs[k] >= n_rows) return FDAL_ERR_SHAPE;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return FDAL_ERR_CUDA;
  CUS(cudaSetDevice(device));
  DevBuf<long long> d_rp;
  DevBuf<int> d_ci, d_dofs;
  DevBuf<double> d_v, d_phi, d_w;
  DevBuf<unsigned long long> d_miss;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long");
  CUS(d_rp.upload((const long long *)row_ptr, (size_t)n_rows + 1));
  CUS(d_ci.upload(col, (size_t)nnz));
  CUS(d_v.upload(val_inout, (size_t)nnz));
  CUS(d_dofs.upload(point_dofs, (size_t)(n_points * dofs_per_cell)));
  CUS(d_phi.upload(point_phi, (size_t)(n_points * dofs_per_cell)));
  CUS(d_w.upload(weight, (size_t)n_points));
  CUS(d_miss.alloc(1));
  CUS(cudaMemset(d_miss.p, 0, sizeof(unsigned long long)));
  const long long nt = (long long)n_points * dofs_per_cell;
  if (nt > 0) {
    const int tb = 256;
    k_al_term_scatter<<<(unsigned)((nt + tb - 1) / tb), tb>>>(n_points, dofs_per_cell, d_dofs.p, d_phi.p, d_w.p, d_rp.p,
                                                              d_ci.p, d_v.p, d_miss.p);
    CUS(cudaGetLastError());
  }
  unsigned long long miss = 0;
  CUS(cudaMemcpy(&miss, d_miss.p, sizeof(miss), cudaMemcpyDeviceToHost));
  if (nnz) CUS(cudaMemcpy(val_inout, d_v.p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost));
  if (n_missing_out) *n_missing_out = (int64_t)miss;
  return miss ? FDAL_ERR_SHAPE : FDAL_OK;
}

}  // extern "C"
