// bsr_build.cu — CSR -> BSR conversion on the device (setup phase of fdal_finalize).
//
// The dim-blocked matrices (velocity / elasticity block and the finest AMG operator: 1.8 G scalar entries each on
// the headline problem) are uploaded as scalar CSR anyway; converting them on the host costs two sorts per block
// row on the CPU cores plus a second 15 GB host->device copy.  Here the conversion runs where the data already is:
//
//   k_bsr_count   one warp per block row: the block columns of its b CSR rows (one contiguous entry range) go
//                 into a per-warp hash set in shared memory; the number of distinct ones is the row's block count
//   k_scan_counts one CTA: running 64-bit prefix sum over the counts -> block-row pointers + the total
//   k_bsr_fill    one warp per block row: rebuild the set, extract and bitonic-sort it (ascending block columns,
//                 the order the host conversion produces), then every scalar entry binary-searches its block and
//                 adds its value into the zero-initialised b x b slot
//
// Everything is bound by reading ci (twice) and v (once) from HBM.  The result equals the host conversion
// (csrc/host_finalize.h: host_bsr_convert) bit for bit; duplicate (row, column) entries are summed by atomicAdd
// (two duplicates: order-independent; three or more could differ in the last bit from the host's entry order).
#include "bsr_build.h"

#include <limits.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../../include/fdal.h"

namespace fdal {
namespace {

constexpr int kWarps = 4;  // warps (= block rows in flight) per CTA; the host drops to 2 or 1 for very long block rows
constexpr int kThreads = kWarps * 32;
constexpr int kEmpty = -1;

// insert J into the open-addressing set t[0, H); returns 1 when J was not there yet
__device__ __forceinline__ int set_insert(int *t, int H, int logH, int J) {
  unsigned s = ((unsigned)J * 2654435761u) >> (32 - logH);
  for (int probe = 0; probe < H; ++probe) {  // H >= 2 x the entries of the block row: never exhausted
    const int old = atomicCAS(&t[s], kEmpty, J);
    if (old == kEmpty) return 1;
    if (old == J) return 0;
    s = (s + 1) & (unsigned)(H - 1);
  }
  return 0;
}

template <int B>
__global__ void __launch_bounds__(kThreads) k_bsr_count(int nbr, const int *__restrict__ rp, const int *__restrict__ ci,
                                                         int H, int logH, int *__restrict__ counts) {
  extern __shared__ int smem_i[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int *t = smem_i + (size_t)w * H;
  for (int I = blockIdx.x * nw + w; I < nbr; I += gridDim.x * nw) {
    for (int s = lane; s < H; s += 32) t[s] = kEmpty;
    __syncwarp();
    const int k0 = __ldg(rp + (size_t)I * B), k1 = __ldg(rp + (size_t)I * B + B);
    int cnt = 0;
    for (int k = k0 + lane; k < k1; k += 32) cnt += set_insert(t, H, logH, __ldg(ci + k) / B);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[I] = cnt;
    __syncwarp();
  }
}

// brp[0] = 0, brp[i + 1] = counts[0] + ... + counts[i] (clamped to INT_MAX), *total = the 64-bit sum.  One CTA.
__global__ void __launch_bounds__(1024) k_scan_counts(int n, const int *__restrict__ counts, int *__restrict__ brp,
                                                       long long *total) {
  __shared__ long long wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  long long carry = 0;
  if (threadIdx.x == 0) brp[0] = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + (int)threadIdx.x;
    long long x = i < n ? (long long)counts[i] : 0ll;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[w] = x;
    __syncthreads();
    if (w == 0) {
      long long s = wsum[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      wsum[lane] = s;
    }
    __syncthreads();
    const long long incl = x + (w > 0 ? wsum[w - 1] : 0ll) + carry;
    if (i < n) brp[i + 1] = (int)(incl < (long long)INT_MAX ? incl : (long long)INT_MAX);
    carry += wsum[31];
    __syncthreads();  // wsum is rewritten by the next tile
  }
  if (threadIdx.x == 0) *total = carry;
}

template <int B>
__global__ void __launch_bounds__(kThreads) k_bsr_fill(int nbr, const int *__restrict__ rp, const int *__restrict__ ci,
                                                        const double *__restrict__ v, int H, int logH,
                                                        const int *__restrict__ brp, int *__restrict__ bcj,
                                                        double *bv) {
  extern __shared__ int smem_i[];
  __shared__ int nlist[kWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  int *t = smem_i + (size_t)w * (H + H / 2);
  int *lst = t + H;  // [H / 2] >= distinct block columns of any block row (H >= 2 x its scalar entries)
  for (int I = blockIdx.x * nw + w; I < nbr; I += gridDim.x * nw) {
    for (int s = lane; s < H; s += 32) t[s] = kEmpty;
    if (lane == 0) nlist[w] = 0;
    __syncwarp();
    const int k0 = __ldg(rp + (size_t)I * B), k1 = __ldg(rp + (size_t)I * B + B);
    for (int k = k0 + lane; k < k1; k += 32) set_insert(t, H, logH, __ldg(ci + k) / B);
    __syncwarp();
    for (int s = lane; s < H; s += 32) {
      const int J = t[s];
      if (J != kEmpty) lst[atomicAdd(&nlist[w], 1)] = J;
    }
    __syncwarp();
    const int cnt = nlist[w];
    int P = 1;
    while (P < cnt) P <<= 1;
    for (int s = cnt + lane; s < P; s += 32) lst[s] = INT_MAX;
    __syncwarp();
    for (int kk = 2; kk <= P; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < P; i += 32) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const int a = lst[i], c = lst[ixj];
            if ((a > c) == ((i & kk) == 0)) {
              lst[i] = c;
              lst[ixj] = a;
            }
          }
        }
        __syncwarp();
      }
    const int b0 = __ldg(brp + I);
    for (int s = lane; s < cnt; s += 32) bcj[(size_t)b0 + s] = lst[s];
    const int r1 = __ldg(rp + (size_t)I * B + 1);
    const int r2 = B > 2 ? __ldg(rp + (size_t)I * B + 2) : INT_MAX;
    for (int k = k0 + lane; k < k1; k += 32) {
      const int c = __ldg(ci + k);
      const int J = c / B, q = c - J * B;
      const int r = (k >= r1 ? 1 : 0) + (k >= r2 ? 1 : 0);
      int lo = 0, hi = cnt;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] < J)
          lo = mid + 1;
        else
          hi = mid;
      }
      atomicAdd(bv + ((size_t)b0 + lo) * (B * B) + r * B + q, __ldg(v + k));
    }
    __syncwarp();
  }
}

template <class T>
struct Scoped {  // frees on scope exit unless released
  T *p = nullptr;
  ~Scoped() {
    if (p) cudaFree(p);
  }
  T *release() {
    T *q = p;
    p = nullptr;
    return q;
  }
};

}  // namespace

#define BCU(call)                                                                                      \
  do {                                                                                                 \
    const cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess) {                                                                           \
      fprintf(stderr, "fdal bsr_build: %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      cudaGetLastError();                                                                              \
      return BSR_BUILD_CUDA_ERROR;                                                                     \
    }                                                                                                  \
  } while (0)

int bsr_from_csr_device(cudaStream_t stream, int sms, int nr, long long nnz, const int *rp, const int *ci, const double *v,
                        int b, int max_row_entries, double max_fill, BsrBuilt *out) {
  *out = BsrBuilt();
  if (b < 2 || b > 3 || nr <= 0 || nr % b || nnz <= 0 || max_row_entries <= 0) return BSR_BUILD_DECLINED;
  const int nbr = nr / b;
  int H = 32, logH = 5;
  while (H < 2 * max_row_entries && H < (1 << 20)) {
    H <<= 1;
    ++logH;
  }
  int warps = kWarps;
  size_t smem_count = 0, smem_fill = 0;
  int dev = 0, optin = 0;
  BCU(cudaGetDevice(&dev));
  BCU(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // dynamic shared memory above 48 KB is opt-in per kernel; the limit is the device maximum minus the kernel's
  // static shared memory (k_bsr_fill keeps 16 bytes of it), so ask for exactly what the launch uses
  const void *f_count = b == 2 ? (const void *)k_bsr_count<2> : (const void *)k_bsr_count<3>;
  const void *f_fill = b == 2 ? (const void *)k_bsr_fill<2> : (const void *)k_bsr_fill<3>;
  cudaFuncAttributes fa_count, fa_fill;
  BCU(cudaFuncGetAttributes(&fa_count, f_count));
  BCU(cudaFuncGetAttributes(&fa_fill, f_fill));
  if (H < 2 * max_row_entries) return BSR_BUILD_DECLINED;
  for (;; warps >>= 1) {  // fewer block rows in flight per CTA when one block row's set is large
    smem_count = (size_t)warps * H * sizeof(int);
    smem_fill = (size_t)warps * (H + H / 2) * sizeof(int);
    if (smem_fill + fa_fill.sharedSizeBytes <= (size_t)optin && smem_count + fa_count.sharedSizeBytes <= (size_t)optin) break;
    if (warps == 1) return BSR_BUILD_DECLINED;
  }
  const int threads = warps * 32;
  if (smem_count > 48 * 1024) BCU(cudaFuncSetAttribute(f_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
  if (smem_fill > 48 * 1024) BCU(cudaFuncSetAttribute(f_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fill));
  Scoped<int> counts, brp, bcj;
  Scoped<long long> total;
  Scoped<double> bv;
  BCU(cudaMalloc((void **)&counts.p, (size_t)nbr * sizeof(int)));
  BCU(cudaMalloc((void **)&brp.p, ((size_t)nbr + 1) * sizeof(int)));
  BCU(cudaMalloc((void **)&total.p, sizeof(long long)));
  // grid: every SM full of CTAs (shared memory is the limit), grid-stride over the block rows
  const int per_sm_count = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)optin / std::max<size_t>(smem_count, 1)));
  const int per_sm_fill = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)optin / std::max<size_t>(smem_fill, 1)));
  const int need = (nbr + warps - 1) / warps;
  const int g_count = std::max(1, std::min(need, sms * per_sm_count));
  const int g_fill = std::max(1, std::min(need, sms * per_sm_fill));
  if (b == 2)
    k_bsr_count<2><<<g_count, threads, smem_count, stream>>>(nbr, rp, ci, H, logH, counts.p);
  else
    k_bsr_count<3><<<g_count, threads, smem_count, stream>>>(nbr, rp, ci, H, logH, counts.p);
  BCU(cudaGetLastError());
  k_scan_counts<<<1, 1024, 0, stream>>>(nbr, counts.p, brp.p, total.p);
  BCU(cudaGetLastError());
  long long nblk = 0;
  BCU(cudaMemcpyAsync(&nblk, total.p, sizeof(long long), cudaMemcpyDeviceToHost, stream));
  BCU(cudaStreamSynchronize(stream));
  if (nblk <= 0 || nblk >= (long long)INT_MAX / (b * b)) return BSR_BUILD_DECLINED;
  if ((double)nblk * b * b > max_fill * (double)nnz) return BSR_BUILD_DECLINED;  // too much zero fill: stay scalar
  BCU(cudaMalloc((void **)&bcj.p, (size_t)nblk * sizeof(int)));
  BCU(cudaMalloc((void **)&bv.p, (size_t)nblk * b * b * sizeof(double)));
  BCU(cudaMemsetAsync(bv.p, 0, (size_t)nblk * b * b * sizeof(double), stream));
  if (b == 2)
    k_bsr_fill<2><<<g_fill, threads, smem_fill, stream>>>(nbr, rp, ci, v, H, logH, brp.p, bcj.p, bv.p);
  else
    k_bsr_fill<3><<<g_fill, threads, smem_fill, stream>>>(nbr, rp, ci, v, H, logH, brp.p, bcj.p, bv.p);
  BCU(cudaGetLastError());
  BCU(cudaStreamSynchronize(stream));
  out->brp = brp.release();
  out->bcj = bcj.release();
  out->bv = bv.release();
  out->nblk = nblk;
  return BSR_BUILD_OK;
}

}  // namespace fdal

// ---- stand-alone entry point (host arrays in and out): the conversion fdal_finalize runs, for tests and tools
extern "C" int64_t fdal_csr_to_bsr(int device, int64_t n_rows, int64_t nnz, const int64_t *row_ptr, const int32_t *col,
                                   const double *val, int32_t block_size, double max_fill, int32_t *brow_ptr_out,
                                   int64_t block_capacity, int32_t *bcol_out, double *bval_out) {
  using namespace fdal;
  if (!row_ptr || n_rows <= 0 || nnz <= 0 || !col || !val || !brow_ptr_out || n_rows >= INT_MAX || nnz >= INT_MAX)
    return -FDAL_ERR_INVALID;
  if (block_size < 2 || block_size > 3 || n_rows % block_size) return -FDAL_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -FDAL_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return -FDAL_ERR_CUDA;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  std::vector<int> rp32((size_t)n_rows + 1);
  int max_entries = 0;
  for (int64_t i = 0; i <= n_rows; ++i) rp32[(size_t)i] = (int)row_ptr[i];
  for (int64_t I = 0; I < n_rows / block_size; ++I)
    max_entries = std::max(max_entries, rp32[(size_t)((I + 1) * block_size)] - rp32[(size_t)(I * block_size)]);
  Scoped<int> d_rp, d_ci;
  Scoped<double> d_v;
  if (cudaMalloc((void **)&d_rp.p, ((size_t)n_rows + 1) * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&d_ci.p, (size_t)nnz * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&d_v.p, (size_t)nnz * sizeof(double)) != cudaSuccess)
    return -FDAL_ERR_ALLOC;
  if (cudaMemcpy(d_rp.p, rp32.data(), ((size_t)n_rows + 1) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_ci.p, col, (size_t)nnz * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_v.p, val, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)
    return -FDAL_ERR_CUDA;
  BsrBuilt built;
  const int st = bsr_from_csr_device(nullptr, sms, (int)n_rows, nnz, d_rp.p, d_ci.p, d_v.p, block_size, max_entries,
                                     max_fill, &built);
  if (st == BSR_BUILD_DECLINED) return -1000;  // not blocked (zero fill / size): no error
  if (st != BSR_BUILD_OK) return -FDAL_ERR_CUDA;
  Scoped<int> o_rp, o_cj;
  Scoped<double> o_v;
  o_rp.p = built.brp;
  o_cj.p = built.bcj;
  o_v.p = built.bv;
  const int64_t nbr = n_rows / block_size;
  if (cudaMemcpy(brow_ptr_out, built.brp, ((size_t)nbr + 1) * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -FDAL_ERR_CUDA;
  if (bcol_out && bval_out && block_capacity >= built.nblk) {
    const size_t bb = (size_t)block_size * block_size;
    if (cudaMemcpy(bcol_out, built.bcj, (size_t)built.nblk * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(bval_out, built.bv, (size_t)built.nblk * bb * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -FDAL_ERR_CUDA;
  }
  return (int64_t)built.nblk;
}
