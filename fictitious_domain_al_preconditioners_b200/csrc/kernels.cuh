// kernels.cuh — hand-written sm_100a kernels of the AL solve path.
//
// Everything here is FP64 on CUDA cores and HBM-bound (arithmetic intensity of a
// CSR SpMV is ~0.17 flop/B), so the design rules are: coalesced streaming of the
// matrix arrays, x gathered through the read-only path (the whole x vector of the
// largest configuration fits the 126 MB L2), epilogues fused into the row pass so
// no intermediate vector goes back to HBM, reductions finished in-kernel
// (last-block pattern, fixed summation order => run-to-run deterministic), and
// grids sized as a multiple of the SM count with grid-stride loops.
//
// Reference call sites these kernels replace: K1 SparseMatrix::vmult,
// K2 coupling_matrix.Tvmult/vmult, K3 the LinearOperator expression
// `K + gamma * Ct * invW * C` (immersed_laplace.cc:884), K4 DiagonalMatrix::vmult,
// K8 PreconditionAMG::vmult (Chebyshev steps, residual, restriction, prolongation),
// K9/K10/K11 the Vector BLAS-1 inside SolverCG / SolverFGMRES.
#pragma once
#include <type_traits>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdal {

constexpr int kBlock = 256;       // threads per CTA for every kernel here
constexpr int kMaxGridPerSM = 8;  // CTAs per SM used to size grid-stride grids
constexpr int kMaxPartials = 148 * 16;
constexpr int kMultiDotGroup = 8;  // basis vectors accumulated per pass in V^T w

struct CsrDev {
  int nrows = 0, ncols = 0;
  long long nnz = 0;
  const int *rp = nullptr;
  const int *ci = nullptr;
  const double *v = nullptr;
  int tpr = 8;  // threads per row
  // as the SECOND matrix of a fused two-matrix pass (Ct beside A): one flag per row chunk of the
  // first matrix, 0 = none of the chunk's rows has an entry here (almost every chunk of Ct: the
  // immersed body touches a thin band of background rows), so the pass skips this matrix's row
  // pointers for the whole chunk.  nullptr: look at every row.
  const unsigned char *chunk_any = nullptr;
};

// x vector with an optional halo part (multi-GPU): columns [0,n_owned) are local,
// [n_owned, ...) index the halo receive buffer of the matrix.
//
// Multi-GPU exchange channels (one process per GPU, peers mapped with cudaIpc over NVLink):
// a channel is a local receive buffer with two slots (epoch parity), one flag word per
// source rank and a device-resident epoch counter.  A sender stores its entries DIRECTLY
// into the receivers' slot (e & 1) with ordinary global stores through the peer mapping,
// fences at system scope and then releases flag[me] = e on every neighbour; a consumer
// acquires flag[q] >= e for all its neighbours before it touches the slot.  The neighbour
// relation is symmetric by construction and every rank issues the same sequence of
// exchanges per channel, so "q released epoch e+1" implies "q is done reading epoch e":
// two slots suffice and no acknowledgement traffic is needed.  Nothing here depends on a
// host-supplied value, so the kernels replay unchanged inside CUDA graphs.
struct ChanDev {
  double *recv = nullptr;                 // local [2][cap]
  unsigned long long *flags = nullptr;    // local [nranks], written by the peers
  unsigned long long *epoch = nullptr;    // local device scalar: last completed push
  unsigned int *counter = nullptr;        // last-block election of the push kernel
  int cap = 0;                            // doubles per slot
  int nnb = 0;                            // neighbours (for broadcast channels: all ranks incl. self)
  const int *nb_rank = nullptr;           // [nnb]
  double *const *nb_recv = nullptr;       // [nnb] peer slot-0 address where MY entries start
  const int *nb_cap = nullptr;            // [nnb] the peer's slot stride
  unsigned long long *const *nb_flag = nullptr;  // [nnb] &peer.flags[me]
  const int *nb_begin = nullptr;          // [nnb + 1] my send-list range per neighbour (gather channels)
};
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// all threads of the block call this; returns once every neighbour has released epoch >= e
__device__ __forceinline__ void chan_wait(const ChanDev &ch, unsigned long long e) {
  for (int k = threadIdx.x; k < ch.nnb; k += blockDim.x) {
    const unsigned long long *f = ch.flags + ch.nb_rank[k];
    while (ld_acquire_sys(f) < e) __nanosleep(20);
  }
  __syncthreads();
}

struct XVec {
  const double *x;
  const double *halo;
  int n_owned;
  // multi-GPU with peer channels: halo = ch->recv + (epoch & 1) * cap.  The row chunks are listed in
  // `order` (the n_interior chunks without halo columns first).  The grid is split in proportion to the
  // work: the first g_bnd CTAs acquire the halo once, before their loop, and share the boundary chunks;
  // the other CTAs share the interior chunks and never wait, so the peers' pushes overlap the interior
  // rows.  The boundary CTAs come FIRST in block order: a grid larger than what is resident starts its
  // last blocks only when earlier ones retire, and boundary blocks started last would run alone at the
  // tail (3-D slabs: 16 % of the rows touch the halo).  No barrier sits inside the row loop (a barrier
  // there stops ptxas from unrolling the row walk: -13 % on the 3-D V-cycle).
  // ch == nullptr: single GPU or NCCL-filled halo buffer.
  const ChanDev *ch = nullptr;
  const int *order = nullptr;
  int n_interior = 0;
  int g_bnd = 0;
};
// ONE load instruction for owned and halo entries (a select on the address, no branch), through the
// read-only path like every other operand of the row walk.  The halo slot is filled by the peers
// before the flag this block acquires in chunk_range(), no thread of the kernel touches a halo address
// before that acquire (interior chunks have no halo columns), and L1 does not survive kernel
// boundaries — so no stale line can be resident when the first ld.global.nc of a halo entry issues.
__device__ __forceinline__ double xload(const XVec &X, int c) {
  const double *p = c < X.n_owned ? X.x + c : X.halo + (c - X.n_owned);
  return __ldg(p);
}
// chunk positions [p, pend) with stride `stride` of this CTA; boundary CTAs acquire the halo here
struct ChunkRange {
  int p, pend, stride;
};
template <bool DIST>
__device__ __forceinline__ ChunkRange chunk_range(XVec &X, int nchunks) {
  if (!DIST) return ChunkRange{(int)blockIdx.x, nchunks, (int)gridDim.x};
  const unsigned long long e = *X.ch->epoch;  // written by the push kernel that precedes this kernel in stream order
  X.halo = X.ch->recv + (size_t)(e & 1ull) * X.ch->cap;
  if ((int)blockIdx.x >= X.g_bnd) return ChunkRange{(int)blockIdx.x - X.g_bnd, X.n_interior, (int)gridDim.x - X.g_bnd};
  chan_wait(*X.ch, e);
  return ChunkRange{X.n_interior + (int)blockIdx.x, nchunks, X.g_bnd};
}
template <bool DIST>
__device__ __forceinline__ int chunk_at(const XVec &X, int p) {
  return DIST ? __ldg(X.order + p) : p;
}

// ---- in-kernel reduction: block partial -> last block sums partials in fixed order
constexpr int kArCap = 128;  // doubles per rank and slot in the scalar all-reduce channel
struct Reducer {
  double *partials;       // [gridDim.x * n_out]
  unsigned int *counter;  // self-resetting (atomicInc wrap)
  double *out;            // [n_out] device scalars
  // multi-GPU: the block that finishes the reduction also sums out[] over the ranks through
  // the scalar all-reduce channel (peer stores over NVLink), so a dot product and its
  // all-reduce are ONE kernel.  nullptr on a single GPU.
  const ChanDev *ar = nullptr;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// returns the block total in thread 0
__device__ __forceinline__ double block_sum(double v, double *smem /*[32]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect smem reuse
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    v = lane < (blockDim.x >> 5) ? smem[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}
// in-place all-reduce (sum) of n <= kArCap block-local values over the ranks; called by ALL
// threads of ONE block per rank.  Every rank stores its values into region [me] of every
// rank's slot (itself included), releases its flag, acquires everybody's flag and adds the
// regions in rank order: the result is bit-identical on every rank.
__device__ __forceinline__ void ar_scalar(const ChanDev &ch, const double *s_vals, int n, double *out) {
  __syncthreads();
  const unsigned long long e = *ch.epoch + 1;
  const size_t slot = (size_t)(e & 1ull);
  for (int idx = threadIdx.x; idx < ch.nnb * n; idx += blockDim.x) {
    const int k = idx / n, o = idx - k * n;
    ch.nb_recv[k][slot * ch.nb_cap[k] + o] = s_vals[o];
  }
  __threadfence_system();
  __syncthreads();
  for (int k = threadIdx.x; k < ch.nnb; k += blockDim.x) st_release_sys(ch.nb_flag[k], e);
  chan_wait(ch, e);
  const double *base = ch.recv + slot * ch.cap;
  for (int o = threadIdx.x; o < n; o += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < ch.nnb; ++r) a += base[(size_t)r * kArCap + o];
    out[o] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) *ch.epoch = e;
}
// Every thread of every block must call this, once per slot and in slot order.
// Deterministic: the final summation order depends only on gridDim.x and
// blockDim.x.  Returns true (for all threads of the block) in the one block that
// finished the reduction, after R.out[] has been written.
__device__ __forceinline__ bool reduce_finalize(double v, const Reducer &R, int slot, int n_out, double *smem) {
  const double tot = block_sum(v, smem);
  if (threadIdx.x == 0) R.partials[(size_t)blockIdx.x * n_out + slot] = tot;
  if (slot != n_out - 1) return false;
  // last slot of this block: publish all its partials and elect the last block
  __shared__ bool s_last;
  __shared__ double s_res[kArCap];
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicInc(R.counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  const bool last = s_last;
  if (last) {
    __threadfence();
    for (int o = 0; o < n_out; ++o) {
      double a = 0.0;
      for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) a += R.partials[(size_t)b * n_out + o];
      a = block_sum(a, smem);
      if (threadIdx.x == 0) {
        if (R.ar)
          s_res[o] = a;
        else
          R.out[o] = a;
      }
    }
    if (R.ar) ar_scalar(*R.ar, s_res, n_out, R.out);
  }
  return last;
}

// ---- SpMV epilogues: called by the first lane of the row group with the row sum.
// operator() returns the row's contribution to a fused reduction (0 if none).
struct EpiAssign {  // y = alpha * A x
  double *y;
  double alpha;
  static constexpr bool kReduce = false;
  __device__ double operator()(int i, double s) const {
    y[i] = alpha * s;
    return 0.0;
  }
};
struct EpiAdd {  // y += alpha * A x
  double *y;
  double alpha;
  static constexpr bool kReduce = false;
  __device__ double operator()(int i, double s) const {
    y[i] += alpha * s;
    return 0.0;
  }
  struct Pre {
    double yi;
  };
  __device__ Pre prefetch(int i) const { return Pre{y[i]}; }
  __device__ double finish(int i, double s, const Pre &p) const {
    y[i] = p.yi + alpha * s;
    return 0.0;
  }
};
struct EpiResid {  // y = b - A x     (also u0 - Ct v1 in the preconditioners)
  double *y;
  const double *b;
  static constexpr bool kReduce = false;
  __device__ double operator()(int i, double s) const {
    y[i] = b[i] - s;
    return 0.0;
  }
  struct Pre {
    double bi;
  };
  __device__ Pre prefetch(int i) const { return Pre{b[i]}; }
  __device__ double finish(int i, double s, const Pre &p) const {
    y[i] = p.bi - s;
    return 0.0;
  }
};
// phase 1 of the augmented operator on the C rows (m of them):
//   w = C x (+ sub: w -= sub[i]) ; y1 = w (optional) ; t = a * winv .* w + add (optional)
// winv == nullptr -> t = a * w + add (exact mass inverse applied afterwards)
struct EpiCouple {
  double *t;
  const double *winv;
  double a;
  const double *add;
  double *y1;
  static constexpr bool kReduce = false;
  __device__ double operator()(int i, double s) const {
    if (y1) y1[i] = s;
    double tv = a * (winv ? winv[i] * s : s);
    if (add) tv += add[i];
    t[i] = tv;
    return 0.0;
  }
};
struct EpiAddDotX {  // y += A x ; reduce x_i * y_i  (joins the overlapped augmented apply)
  double *y;
  const double *x;
  static constexpr bool kReduce = true;
  __device__ double operator()(int i, double s) const {
    const double yn = y[i] + s;
    y[i] = yn;
    return x[i] * yn;
  }
  struct Pre {
    double yi, xi;
  };
  __device__ Pre prefetch(int i) const { return Pre{y[i], x[i]}; }
  __device__ double finish(int i, double s, const Pre &p) const {
    const double yn = p.yi + s;
    y[i] = yn;
    return p.xi * yn;
  }
};
struct EpiDotX {  // y = A x ; reduce x_i * y_i  (p.Ap of CG)
  double *y;
  const double *x;
  static constexpr bool kReduce = true;
  __device__ double operator()(int i, double s) const {
    y[i] = s;
    return x[i] * s;
  }
  struct Pre {
    double xi;
  };
  __device__ Pre prefetch(int i) const { return Pre{x[i]}; }
  __device__ double finish(int i, double s, const Pre &p) const {
    y[i] = s;
    return p.xi * s;
  }
};
// fused Chebyshev step: r = b - A x ; d = c1 d + c2 D^-1 r ; xout = xin + d
template <bool REDUCE>
struct EpiCheb {
  const double *b, *invd, *xin;
  double *d, *xout;
  double c1, c2;
  int first;  // 1: d is not read (c1 term absent)
  static constexpr bool kReduce = REDUCE;
  __device__ double operator()(int i, double s) const {
    const double bi = b[i];
    const double r = bi - s;
    double dn = c2 * invd[i] * r;
    if (!first) dn += c1 * d[i];
    d[i] = dn;
    const double xo = xin[i] + dn;
    xout[i] = xo;
    return REDUCE ? bi * xo : 0.0;  // r.z of the enclosing CG when b is the CG residual
  }
  // split form for kernels that fetch the row's operands before walking the row (k_bsr_spmv, U >= 4):
  // row i of b, invd, d, xin is only ever written by row i's own epilogue, so the early read is safe
  struct Pre {
    double bi, invdi, di, xi;
  };
  __device__ Pre prefetch(int i) const { return Pre{b[i], invd[i], first ? 0.0 : d[i], xin[i]}; }
  __device__ double finish(int i, double s, const Pre &p) const {
    const double r = p.bi - s;
    double dn = c2 * p.invdi * r;
    if (!first) dn += c1 * p.di;
    d[i] = dn;
    const double xo = p.xi + dn;
    xout[i] = xo;
    return REDUCE ? p.bi * xo : 0.0;
  }
};
// epilogues that offer the split prefetch / finish form (a nested Pre type)
template <class Epi, class = void>
struct EpiHasPrefetch {
  static constexpr bool value = false;
};
template <class Epi>
struct EpiHasPrefetch<Epi, std::void_t<typename Epi::Pre>> {
  static constexpr bool value = true;
};
struct NoPre {};
template <class Epi, bool ON>
struct PreOf {
  using type = NoPre;
};
template <class Epi>
struct PreOf<Epi, true> {
  using type = typename Epi::Pre;
};

// ---- CSR SpMV, TPR threads per row, grid-stride over row chunks ----------------
// row-group partial sum with U independent (value, column, x) load chains in flight per
// lane: all U value/column loads are issued back to back (predicated, no divergence),
// then the U gathers, then the FMAs — memory-level parallelism no longer depends on the
// row being long enough to fill an unrolled trip.
template <int TPR, int U>
__device__ __forceinline__ double row_partial(const CsrDev &A, const XVec &X, int k0, int k1, int lane) {
  double s = 0.0;
  if (U > 1) {
    for (int k = k0 + lane; k < k1; k += U * TPR) {
      double vv[U];
      int cc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int kk = k + u * TPR;
        const bool ok = kk < k1;
        vv[u] = ok ? __ldg(A.v + kk) : 0.0;
        cc[u] = ok ? __ldg(A.ci + kk) : -1;
      }
      double xx[U];
#pragma unroll
      for (int u = 0; u < U; ++u) xx[u] = cc[u] >= 0 ? xload(X, cc[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < U; ++u) s += vv[u] * xx[u];
    }
  } else {
    for (int k = k0 + lane; k < k1; k += TPR) s += __ldg(A.v + k) * xload(X, __ldg(A.ci + k));
  }
  return s;
}

template <int TPR, class Epi, int U = 1, bool DIST = false>
__global__ void __launch_bounds__(kBlock) k_spmv(CsrDev A, XVec X, Epi epi, Reducer R) {
  __shared__ double smem[32];
  constexpr int rows_per_block = kBlock / TPR;
  const int lane = threadIdx.x & (TPR - 1);
  const int local_row = threadIdx.x / TPR;
  double contrib = 0.0;
  const int nchunks = (A.nrows + rows_per_block - 1) / rows_per_block;
  const ChunkRange cr = chunk_range<DIST>(X, nchunks);
  for (int p = cr.p; p < cr.pend; p += cr.stride) {
    const int row = chunk_at<DIST>(X, p) * rows_per_block + local_row;
    double s = 0.0;
    if (row < A.nrows) s = row_partial<TPR, U>(A, X, __ldg(A.rp + row), __ldg(A.rp + row + 1), lane);
#pragma unroll
    for (int o = TPR >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, TPR);
    if (lane == 0 && row < A.nrows) contrib += epi((int)row, s);
  }
  if (Epi::kReduce) reduce_finalize(contrib, R, 0, 1, smem);
}

// ---- fused augmented apply, phase 2:  y = A x + Ct t   (+ fused x.y) -------------
// Two CSR matrices with the same row partition are walked in one row pass, so A x
// never goes to memory before the coupling term is added (K3).  t already carries
// gamma * W^-1 C x (phase 1), and for the block system also + x1.
template <int TPR, class Epi, int U = 1, bool DIST = false>
__global__ void __launch_bounds__(kBlock) k_spmv2(CsrDev A, XVec X, CsrDev Ct, const double *__restrict__ t, Epi epi,
                                                   Reducer R) {
  __shared__ double smem[32];
  constexpr int rows_per_block = kBlock / TPR;
  const int lane = threadIdx.x & (TPR - 1);
  const int local_row = threadIdx.x / TPR;
  double contrib = 0.0;
  const int nchunks = (A.nrows + rows_per_block - 1) / rows_per_block;
  const ChunkRange cr = chunk_range<DIST>(X, nchunks);
  for (int p = cr.p; p < cr.pend; p += cr.stride) {
    const int chunk = chunk_at<DIST>(X, p);
    const int row = chunk * rows_per_block + local_row;
    const bool second = !Ct.chunk_any || __ldg(Ct.chunk_any + chunk);
    double s = 0.0;
    if (row < A.nrows) {
      int c0 = 0, c1 = 0;
      if (second) {
        c0 = __ldg(Ct.rp + row);
        c1 = __ldg(Ct.rp + row + 1);
      }
      s = row_partial<TPR, U>(A, X, __ldg(A.rp + row), __ldg(A.rp + row + 1), lane);
      for (int k = c0 + lane; k < c1; k += TPR) s += __ldg(Ct.v + k) * __ldg(t + __ldg(Ct.ci + k));
    }
#pragma unroll
    for (int o = TPR >> 1; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, TPR);
    if (lane == 0 && row < A.nrows) contrib += epi((int)row, s);
  }
  if (Epi::kReduce) reduce_finalize(contrib, R, 0, 1, smem);
}

// ---- merge-path CSR SpMV (Merrill & Garland, SC'16):  y = alpha A x  or  y += alpha A x ----------
// The merge of the nrows row-end offsets with the nnz non-zero indices is cut into equal pieces of
// kMergeIPT items per thread, whatever the row lengths: a matrix with fewer rows than the GPU has
// row groups (C = Ct^T: one row per multiplier, a few thousand rows) or with very uneven rows
// (restriction operators, constraint rows) still spreads evenly over all SMs, where the row-group
// kernel k_spmv leaves lanes and SMs idle.  A thread walks its piece sequentially — consume a
// non-zero, or finish a row — and hands the unfinished tail of its last row on as a carry.  Carries are
// folded in a fixed order (deterministic, no atomics): inside the CTA through shared memory, across CTAs
// by k_spmv_merge_fixup.  Only linear epilogues (assign / add), because a row's sum is completed after
// the pass.
constexpr int kMergeIPT = 8;                      // merge items per thread
constexpr int kMergeTile = kMergeIPT * kBlock;    // merge items per CTA

__global__ void __launch_bounds__(kBlock) k_spmv_merge(CsrDev A, const double *__restrict__ x, double alpha, int add,
                                                        double *__restrict__ y, int *__restrict__ carry_row,
                                                        double *__restrict__ carry_val) {
  __shared__ int s_row[kBlock];
  __shared__ double s_val[kBlock];
  const long long n_items = (long long)A.nrows + A.nnz;
  const long long d0 = min((long long)blockIdx.x * kMergeTile + (long long)threadIdx.x * kMergeIPT, n_items);
  // split point of diagonal d0: i rows finished, j non-zeros consumed, i + j = d0
  long long lo = max(0ll, d0 - A.nnz), hi = min(d0, (long long)A.nrows);
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if ((long long)__ldg(A.rp + mid + 1) <= d0 - mid - 1)
      lo = mid + 1;
    else
      hi = mid;
  }
  int row = (int)lo;
  long long j = d0 - lo;
  double run = 0.0;
  int row_end = row < A.nrows ? __ldg(A.rp + row + 1) : 0;
#pragma unroll
  for (int s = 0; s < kMergeIPT; ++s) {
    if (d0 + s >= n_items || row >= A.nrows) break;
    if (j < row_end) {
      run += __ldg(A.v + j) * x[__ldg(A.ci + j)];
      ++j;
    } else {
      // this thread finishes the row: its own part now, the earlier threads' carries in the fix-up
      y[row] = add ? y[row] + alpha * run : alpha * run;
      run = 0.0;
      ++row;
      row_end = row < A.nrows ? __ldg(A.rp + row + 1) : 0;
    }
  }
  s_row[threadIdx.x] = row;
  s_val[threadIdx.x] = run;
  __syncthreads();
  // carries with the same row form a run of consecutive threads: its last thread folds the run
  const bool last_of_run = threadIdx.x == kBlock - 1 || s_row[threadIdx.x + 1] != row;
  if (last_of_run && row < A.nrows) {
    double tot = 0.0;
    int t = threadIdx.x;
    while (t >= 0 && s_row[t] == row) --t;
    for (++t; t <= (int)threadIdx.x; ++t) tot += s_val[t];  // ascending thread order: fixed summation order
    if (threadIdx.x == kBlock - 1) {
      carry_row[blockIdx.x] = row;  // the row continues in the next CTA (or ends exactly here: then a later
      carry_val[blockIdx.x] = tot;  // thread has written / will write its tail and the fix-up adds this)
    } else {
      y[row] += alpha * tot;  // the row was finished by thread threadIdx.x + 1 of this CTA (visible after the barrier)
    }
  } else if (threadIdx.x == kBlock - 1) {
    carry_row[blockIdx.x] = A.nrows;  // nothing to carry
    carry_val[blockIdx.x] = 0.0;
  }
}
// fold the CTA carries: CTAs whose carry belongs to the same row are consecutive
__global__ void __launch_bounds__(kBlock) k_spmv_merge_fixup(int n_tiles, int nrows, double alpha,
                                                              const int *__restrict__ carry_row,
                                                              const double *__restrict__ carry_val,
                                                              double *__restrict__ y) {
  for (int b = blockIdx.x * kBlock + threadIdx.x; b < n_tiles; b += gridDim.x * kBlock) {
    const int row = carry_row[b];
    if (row >= nrows) continue;
    if (b + 1 < n_tiles && carry_row[b + 1] == row) continue;  // not the last CTA of this row's run
    int t = b;
    while (t >= 0 && carry_row[t] == row) --t;
    double tot = 0.0;
    for (++t; t <= b; ++t) tot += carry_val[t];
    y[row] += alpha * tot;
  }
}

// ---- BSR SpMV for node-interleaved vector-valued blocks (Stokes velocity, elasticity) ----
// B x B dense blocks (B = 2, 3) share one column index: (8 B^2 + 4) bytes per block instead
// of 12 B^2 for scalar CSR (-25 % / -30 % HBM traffic on the dominant fine-level matrix).
// Each block's B^2 values are contiguous (AoS): a lane issues the B^2 + B + 1 loads of its
// block back to back (a per-row plane layout was measured 25-40 % slower: it needs every
// cache line to survive in L1 between k-iterations).  Epilogues are the scalar ones, run
// by B lanes in parallel for the B rows of the block row.  TWO adds a scalar CSR matrix with the same scalar
// rows (Ct) in the same pass: y = A x + Ct t.
struct BsrDev {
  int nbrows = 0;
  const int *rp = nullptr;    // [nbrows + 1] block-row pointer
  const int *cj = nullptr;    // [nblocks] block column
  const double *v = nullptr;  // per block row: B*B planes of (rp[I+1]-rp[I]) doubles (or AoS blocks)
  int tpr = 8;
  int aos = 0;
};

// 3x3 blocks, one block per lane in flight: 8 CTAs/SM (32 registers) for every epilogue.  Without the bound the
// epilogue with the fused reduction (EpiCheb<true>) took 40 registers -> 6 CTAs/SM -> 565 us against 505 us for
// the same step without the reduction (profiles/r2_ncu_full_k_bsr_spmv_stokes3d_nel40.md); no variant spills.
template <int B, int TPR, class Epi, bool TWO, int U = 1, bool DIST = false>
__global__ void __launch_bounds__(kBlock, (B == 3 && U == 1) ? 8 : 0) k_bsr_spmv(BsrDev A, XVec X, CsrDev C2, const double *__restrict__ t2, Epi epi,
                                                      Reducer R) {
  __shared__ double smem[32];
  constexpr int rows_per_block = kBlock / TPR;
  constexpr bool kPre = U >= 4 && EpiHasPrefetch<Epi>::value && TPR >= B;
  const int lane = threadIdx.x & (TPR - 1);
  const int local_row = threadIdx.x / TPR;
  double contrib = 0.0;
  // U >= 4 (opt-in variant): the block-row pointers of the NEXT grid-stride step are fetched while
  // this step's blocks are in flight, which takes one DRAM latency out of every step's chain
  const int nchunks = (A.nbrows + rows_per_block - 1) / rows_per_block;
  const ChunkRange cr = chunk_range<DIST>(X, nchunks);
  int k0n = 0, k1n = 0;
  if constexpr (U >= 4) {
    if (cr.p < cr.pend) {
      const long long I0 = (long long)chunk_at<DIST>(X, cr.p) * rows_per_block + local_row;
      if (I0 < A.nbrows) {
        k0n = __ldg(A.rp + I0);
        k1n = __ldg(A.rp + I0 + 1);
      }
    }
  }
  for (int p = cr.p; p < cr.pend; p += cr.stride) {
    const int chunk = chunk_at<DIST>(X, p);
    const long long I = (long long)chunk * rows_per_block + local_row;
    double s[B];
#pragma unroll
    for (int r = 0; r < B; ++r) s[r] = 0.0;
    [[maybe_unused]] typename PreOf<Epi, kPre>::type pre{};
    if (I < A.nbrows) {
      int k0, k1;
      if constexpr (U >= 4) {
        k0 = k0n;
        k1 = k1n;
        const int pn = p + cr.stride;
        if (pn < cr.pend) {
          const long long In = (long long)chunk_at<DIST>(X, pn) * rows_per_block + local_row;
          if (In < A.nbrows) {
            k0n = __ldg(A.rp + In);
            k1n = __ldg(A.rp + In + 1);
          }
        }
      } else {
        k0 = __ldg(A.rp + I);
        k1 = __ldg(A.rp + I + 1);
      }
      const int nb = k1 - k0;
      const double *vb = A.v + (size_t)k0 * (B * B);
      if constexpr (kPre)
        if (lane < B) pre = epi.prefetch((int)(I * B + lane));
      for (int kk = lane; kk < nb; kk += U * TPR) {
        // U blocks per lane in flight: all loads of the U blocks are issued before any FMA
        double xj[U][B], a[U][B * B];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int k = kk + u * TPR;
          const bool ok = k < nb;
          const int c = (ok ? __ldg(A.cj + k0 + k) : 0) * B;
          const double *xp = c < X.n_owned ? X.x + c : X.halo + (c - X.n_owned);
#pragma unroll
          for (int q = 0; q < B * B; ++q)
            a[u][q] = ok ? __ldg(vb + (size_t)k * (B * B) + q) : 0.0;
#pragma unroll
          for (int q = 0; q < B; ++q) xj[u][q] = ok ? __ldg(xp + q) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int r = 0; r < B; ++r)
#pragma unroll
            for (int q = 0; q < B; ++q) s[r] += a[u][r * B + q] * xj[u][q];
      }
      if (TWO && (!C2.chunk_any || __ldg(C2.chunk_any + chunk))) {
#pragma unroll
        for (int r = 0; r < B; ++r) {
          const long long row = I * B + r;
          const int c0 = __ldg(C2.rp + row), c1 = __ldg(C2.rp + row + 1);
          for (int k = c0 + lane; k < c1; k += TPR) s[r] += __ldg(C2.v + k) * __ldg(t2 + __ldg(C2.ci + k));
        }
      }
    }
    // butterfly: every lane of the row group ends up with the B row sums, then lane r
    // runs the epilogue of scalar row r (B lanes in parallel, consecutive addresses)
#pragma unroll
    for (int r = 0; r < B; ++r)
#pragma unroll
      for (int o = TPR >> 1; o > 0; o >>= 1) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o, TPR);
    if (lane < B && I < A.nbrows) {
      double mine = s[0];
#pragma unroll
      for (int r = 1; r < B; ++r)
        if (lane == r) mine = s[r];
      if constexpr (kPre)
        contrib += epi.finish((int)(I * B + lane), mine, pre);
      else
        contrib += epi((int)(I * B + lane), mine);
    }
  }
  if (Epi::kReduce) reduce_finalize(contrib, R, 0, 1, smem);
}

// ---- dense GEMV for the coarsest AMG level: y = Ainv b (one warp per row) --------
// rows [0, nrows) of the given row-major slab (ld = ncols): multi-GPU ranks apply only
// their rows of the replicated inverse
__global__ void __launch_bounds__(kBlock) k_gemv(int nrows, int ncols, const double *__restrict__ Ainv,
                                                  const double *__restrict__ b, double *__restrict__ y) {
  const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nrows) return;
  const double *row = Ainv + (size_t)warp * ncols;
  double s = 0.0;
  for (int k = lane; k < ncols; k += 32) s += __ldg(row + k) * __ldg(b + k);
  s = warp_sum(s);
  if (lane == 0) y[warp] = s;
}
// y = a * (W x) (+ add) with a dense row-major n x n W: the exact W^-1 = M^-1 (or M^-2) of a small
// multiplier space kept resident in L2 (opt-in, FDAL_DENSE_WINV; apply_winv_scaled).  One warp per row.
__global__ void __launch_bounds__(kBlock) k_gemv_axpb(int n, const double *__restrict__ W, const double *__restrict__ x,
                                                       double a, const double *__restrict__ add,
                                                       double *__restrict__ y) {
  const int warp = (blockIdx.x * kBlock + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double *row = W + (size_t)warp * n;
  double s0 = 0.0, s1 = 0.0;
  int k = lane;
  for (; k + 32 < n; k += 64) {
    s0 += __ldg(row + k) * __ldg(x + k);
    s1 += __ldg(row + k + 32) * __ldg(x + k + 32);
  }
  if (k < n) s0 += __ldg(row + k) * __ldg(x + k);
  const double s = warp_sum(s0 + s1);
  if (lane == 0) y[warp] = a * s + (add ? add[warp] : 0.0);
}
// C = A A for a dense row-major n x n A (setup, once: M^-2 from M^-1)
__global__ void k_dense_square(int n, const double *__restrict__ A, double *__restrict__ C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= n) return;
  const double *row = A + (size_t)r * n;
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += row[k] * A[(size_t)k * n + c];
  C[(size_t)r * n + c] = s;
}
// ---- exchange-channel kernels (multi-GPU, peers mapped through cudaIpc) ------------------
// elect the last block of a push kernel; it releases the new epoch to every neighbour
__device__ __forceinline__ void chan_publish(const ChanDev &ch, unsigned long long e) {
  __shared__ bool s_last_push;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int nblk = gridDim.x * gridDim.y;
    s_last_push = atomicInc(ch.counter, nblk - 1) == nblk - 1;
  }
  __syncthreads();
  if (s_last_push) {
    __threadfence_system();
    for (int k = threadIdx.x; k < ch.nnb; k += blockDim.x) st_release_sys(ch.nb_flag[k], e);
    if (threadIdx.x == 0) *ch.epoch = e;
  }
}
// halo push: entry i of my send list (grouped by neighbour) goes straight into that neighbour's
// receive slot — the "pack" and the transfer are the same NVLink stores
__global__ void __launch_bounds__(kBlock) k_chan_push_gather(ChanDev ch, const int *__restrict__ idx,
                                                              const double *__restrict__ x) {
  const unsigned long long e = *ch.epoch + 1;
  chan_wait(ch, e - 1);  // every neighbour has finished reading slot (e & 1) of epoch e - 2
  const size_t slot = (size_t)(e & 1ull);
  const int n = ch.nb_begin[ch.nnb];
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    int k = 0;
    while (i >= ch.nb_begin[k + 1]) ++k;
    ch.nb_recv[k][slot * ch.nb_cap[k] + (i - ch.nb_begin[k])] = x[idx[i]];
  }
  chan_publish(ch, e);
}
// broadcast push (all-gather / vector all-reduce): x[0:n) into my region of every rank (blockIdx.y)
__global__ void __launch_bounds__(kBlock) k_chan_push_bcast(ChanDev ch, const double *__restrict__ x, int n) {
  const unsigned long long e = *ch.epoch + 1;
  chan_wait(ch, e - 1);
  const size_t slot = (size_t)(e & 1ull);
  const int k = blockIdx.y;
  double *dst = ch.nb_recv[k] + slot * ch.nb_cap[k];
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) dst[i] = x[i];
  chan_publish(ch, e);
}
// consumer of a broadcast channel.  SUM: out[i] = sum over ranks (in rank order: bit-identical on
// all ranks) of region r; else the regions tile the slot (all-gather) and out = the slot.
// scale_in/add fuse the epilogue of the coupling phase 1 (see k_couple_elem).
template <bool SUM>
__global__ void __launch_bounds__(kBlock) k_chan_collect(ChanDev ch, double *__restrict__ out, int n, int stride) {
  const unsigned long long e = *ch.epoch;
  chan_wait(ch, e);
  const double *base = ch.recv + (size_t)(e & 1ull) * ch.cap;
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    if (SUM) {
      double a = 0.0;
      for (int r = 0; r < ch.nnb; ++r) a += base[(size_t)r * stride + i];
      out[i] = a;
    } else {
      out[i] = base[i];
    }
  }
}
// halo pack: buf[i] = x[idx[i]]
__global__ void __launch_bounds__(kBlock) k_pack(int n, const int *__restrict__ idx, const double *__restrict__ x,
                                                  double *__restrict__ buf) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) buf[i] = x[idx[i]];
}
// the EpiCouple epilogue as a stand-alone pass (multi-GPU: after the all-reduce of C x)
__global__ void __launch_bounds__(kBlock) k_couple_elem(int n, const double *__restrict__ w,
                                                         const double *__restrict__ winv, double a,
                                                         const double *__restrict__ add, double *y1, double *t) {
  for (int i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const double s = w[i];
    if (y1) y1[i] = s;
    double tv = a * (winv ? winv[i] * s : s);
    if (add) tv += add[i];
    t[i] = tv;
  }
}

// ---- Krylov vector kernels (K9-K11) ------------------------------------------------
// device scalar slots of one CG instance
enum { S_RHO = 0, S_RHO_OLD = 1, S_PV = 2, S_RR = 3, S_COUNT = 8 };

__global__ void __launch_bounds__(kBlock) k_dot(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                                 Reducer R) {
  __shared__ double smem[32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    s += a[i] * b[i];
  reduce_finalize(s, R, 0, 1, smem);
}
// z = d .* r and r.z in one pass (Jacobi preconditioner of the mass-matrix CG)
__global__ void __launch_bounds__(kBlock) k_diag_prec_dot(long long n, const double *__restrict__ d,
                                                           const double *__restrict__ r, double *__restrict__ z,
                                                           Reducer R) {
  __shared__ double smem[32];
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const double zi = d[i] * r[i];
    z[i] = zi;
    s += zi * r[i];
  }
  reduce_finalize(s, R, 0, 1, smem);
}
// p = z + beta p, beta = rho / rho_old (rho_old = +inf on the first iteration => p = z;
// rho_old == 0 only when the residual is exactly zero, e.g. a zero right-hand side in
// a fixed-count mass solve: the iterate is frozen instead of producing 0/0)
__global__ void __launch_bounds__(kBlock) k_cg_update_p(long long n, const double *__restrict__ z, double *__restrict__ p,
                                                         const double *__restrict__ scal) {
  const double rho = scal[S_RHO], rho_old = scal[S_RHO_OLD];
  const double beta = (isinf(rho_old) || rho_old == 0.0) ? 0.0 : rho / rho_old;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    p[i] = z[i] + beta * p[i];
}
// x += alpha p ; r -= alpha v ; rr = r.r ; rho_old = rho      (alpha = rho / p.v)
// guard: a zero denominator (exactly converged fixed-count mass solve) freezes the iterate
__global__ void __launch_bounds__(kBlock) k_cg_update_xr(long long n, long long n_dot, const double *__restrict__ p,
                                                          const double *__restrict__ v, double *__restrict__ x,
                                                          double *__restrict__ r, double *scal, Reducer R) {
  __shared__ double smem[32];
  const double rho = scal[S_RHO], pv = scal[S_PV];
  const double alpha = (pv != 0.0) ? rho / pv : 0.0;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * v[i];
    r[i] = ri;
    if (i < n_dot) s += ri * ri;
  }
  // the block that finishes the reduction also rotates rho: every block has read
  // rho before it published its partial, so nobody can still observe the old value
  const bool last = reduce_finalize(s, R, 0, 1, smem);
  if (last && threadIdx.x == 0) scal[S_RHO_OLD] = rho;
}
__global__ void k_set_scalar(double *p, double v) { *p = v; }

// ---- whole Jacobi-PCG mass solve in ONE CTA (K5: the device replacement of
// SparseDirectUMFPACK::vmult for the immersed mass matrix, m <= kMassCtaMaxRows) -----
// The immersed mass matrix is tiny (m = 10^2..10^4 rows, 3-9 non-zeros per row): as
// separate kernels one solve is ~4*its launches of a few hundred nanoseconds of work
// each.  Here one 1024-thread CTA runs the calibrated fixed number of iterations with
// __syncthreads() between phases; vectors stay in L1/L2.  `repeat` = 2 applies M^-1
// twice (W^-1 = M^-1 M^-1, immersed_laplace.cc:875-876).  y = a * result (+ add).
constexpr int kMassCtaThreads = 1024;
constexpr int kMassCtaMaxRows = 16384;
constexpr int kMassCtaSmemRows = 6144;  // 4 vectors of m doubles in shared memory up to here

// block-wide sum broadcast to every thread with ONE __syncthreads: warp partials go to one
// of two alternating shared buffers, then every warp re-reduces all partials redundantly
// (blockDim.x multiple of 32, <= 1024).  `buf` must alternate between consecutive calls.
__device__ __forceinline__ double block_allsum(double v, double *smem /*[2][32]*/, int buf) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double *s = smem + buf * 32;
  v = warp_sum(v);
  if (lane == 0) s[w] = v;
  __syncthreads();
  double t = lane < (int)(blockDim.x >> 5) ? s[lane] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// SMEM: the four CG vectors live in dynamic shared memory (4 * m doubles), else in `ws`.
// Every thread owns the rows i = tid, tid + nt, ...: only the mat-vec reads other threads'
// entries (of p), so one iteration needs three barriers (two inside the reductions).
template <bool SMEM>
__global__ void __launch_bounds__(kMassCtaThreads) k_mass_pcg_cta(CsrDev M, const double *__restrict__ invd, int its,
                                                                   int repeat, double a, const double *__restrict__ b,
                                                                   const double *__restrict__ add, double *y,
                                                                   double *ws /* 5 * m */) {
  extern __shared__ __align__(16) double svec[];
  __shared__ double red[64];
  const int n = M.nrows;
  double *base = SMEM ? svec : ws;
  double *x = base, *r = base + n, *p = base + 2 * (size_t)n, *v = base + 3 * (size_t)n;
  double *bb = ws + 4 * (size_t)n;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int rep = 0; rep < repeat; ++rep) {
    const double *rhs = rep == 0 ? b : bb;
    double part = 0.0;
    for (int i = tid; i < n; i += nt) {
      const double ri = rhs[i];
      const double zi = __ldg(invd + i) * ri;
      x[i] = 0.0;
      r[i] = ri;
      p[i] = zi;
      part += ri * zi;
    }
    double rho = block_allsum(part, red, 0);  // its barrier also publishes p
    for (int it = 0; it < its; ++it) {
      part = 0.0;
      for (int i = tid; i < n; i += nt) {
        double s = 0.0;
        const int k1 = __ldg(M.rp + i + 1);
        for (int k = __ldg(M.rp + i); k < k1; ++k) s += __ldg(M.v + k) * p[__ldg(M.ci + k)];
        v[i] = s;
        part += p[i] * s;
      }
      const double pv = block_allsum(part, red, 1);
      const double alpha = pv != 0.0 ? rho / pv : 0.0;
      part = 0.0;
      for (int i = tid; i < n; i += nt) {  // own rows only
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * v[i];
        r[i] = ri;
        part += ri * ri * __ldg(invd + i);
      }
      const double rho_new = block_allsum(part, red, 0);  // all mat-vec reads of p are done
      const double beta = rho != 0.0 ? rho_new / rho : 0.0;
      rho = rho_new;
      for (int i = tid; i < n; i += nt) p[i] = __ldg(invd + i) * r[i] + beta * p[i];
      __syncthreads();  // p complete for the next mat-vec
    }
    if (rep + 1 < repeat) {
      for (int i = tid; i < n; i += nt) bb[i] = x[i];  // own rows; re-read by the owner only
    }
  }
  for (int i = tid; i < n; i += nt) y[i] = a * x[i] + (add ? add[i] : 0.0);
}

// ---- whole fixed-count Chebyshev mass solve in ONE persistent multi-CTA kernel (K5 for mass matrices
// too large for one CTA: the co-dimension-0 multiplier space of elliptic_interface, m = 2e4..1e5; the
// 2-D pressure space of stokes_immersed_boundary) --------------------------------------------------
// As separate kernels a Jacobi-PCG mass solve is 4 launches per iteration of ~1 us of work each
// (configs[2]: 244 launches = 0.93 ms per augmented apply, 2.7 % of the HBM roofline).  Chebyshev needs
// no inner products: with the spectral interval of D^-1 M known (Lanczos bounds from the calibration run
// at fdal_finalize) iteration k is  r = b - M x ; d = c1_k d + c2_k D^-1 r ; x += d,  and the only
// coupling between rows is the gather of x.  So: CTA j owns rows [j*rpc, (j+1)*rpc); its slice of M
// (row pointers, columns, values) is staged ONCE into shared memory (the host only picks this kernel
// when every slice fits; larger matrices are bandwidth-bound and take one fused SpMV kernel per
// iteration), d, b, D^-1 and the own part of x stay in shared memory for the whole solve, the iterate is
// double-buffered in global memory (L2-resident), and the CTAs meet at one grid barrier per iteration.
// The grid is at most one CTA per SM, so all CTAs are co-resident and the barrier cannot deadlock
// (`bar` is zeroed by a memset node in front of the launch).  coef[2k], coef[2k+1] = c1_k, c2_k.
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// bar[0]: arrivals (monotonic over the launch), bar[1]: set when a CTA gave up waiting.  The wait is bounded
// (~2 s of SM clocks): should the co-residency assumption ever be violated the kernel ends with a wrong
// result and the flag raised — the host checks it when it verifies the solve at fdal_finalize — instead of
// hanging the device.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();  // cumulative: the block's stores (ordered by the barrier above) before the arrival
    atomicAdd(bar, 1u);
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(bar) < target) {
      if (clock64() - t0 > 4000000000ll) {
        atomicExch(bar + 1, 1u);
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(kBlock) k_mass_cheb_grid(CsrDev M, const double *__restrict__ invd,
                                                            const double *__restrict__ coef, int its, int repeat,
                                                            double a, const double *__restrict__ b,
                                                            const double *__restrict__ add, double *__restrict__ y,
                                                            double *xbuf /* [2][n] */, unsigned int *bar,
                                                            int rpc, int nnz_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = M.nrows;
  const int r0 = min(n, (int)blockIdx.x * rpc), r1 = min(n, r0 + rpc);
  const int nr = r1 - r0;
  double *sd = reinterpret_cast<double *>(smem_raw);
  double *sb = sd + rpc, *sinv = sb + rpc, *sx = sinv + rpc;
  double *sv = sx + rpc;                                  // [nnz_cap]
  int *sci = reinterpret_cast<int *>(sv + nnz_cap);       // [nnz_cap]
  int *srp = sci + nnz_cap;                               // [rpc + 1]
  const int tid = threadIdx.x;
  {
    const int kbase = __ldg(M.rp + r0), kend = __ldg(M.rp + r1);
    for (int k = kbase + tid; k < kend; k += kBlock) {
      sv[k - kbase] = __ldg(M.v + k);
      sci[k - kbase] = __ldg(M.ci + k);
    }
    for (int i = tid; i <= nr; i += kBlock) srp[i] = __ldg(M.rp + r0 + i) - kbase;
  }
  for (int i = tid; i < nr; i += kBlock) sinv[i] = __ldg(invd + r0 + i);
  __syncthreads();
  unsigned int gen = 0;
  for (int rep = 0; rep < repeat; ++rep) {
    int p = 0;
    {  // step 0 from the zero guess: d = c2_0 D^-1 b ; x = d
      const double c2 = __ldg(coef + 1);
      for (int i = tid; i < nr; i += kBlock) {
        const double bi = rep == 0 ? b[r0 + i] : sx[i];  // W = M^2: the second solve's right-hand side is x
        const double dv = c2 * sinv[i] * bi;
        sb[i] = bi;
        sd[i] = dv;
        sx[i] = dv;
        xbuf[r0 + i] = dv;
      }
    }
    for (int k = 1; k < its; ++k) {
      grid_barrier(bar, ++gen * gridDim.x);
      const double c1 = __ldg(coef + 2 * k), c2 = __ldg(coef + 2 * k + 1);
      const double *xin = xbuf + (size_t)p * n;
      double *xout = xbuf + (size_t)(p ^ 1) * n;
      for (int i = tid; i < nr; i += kBlock) {
        double s = 0.0;
        const int k1 = srp[i + 1];
        for (int kk = srp[i]; kk < k1; ++kk) s += sv[kk] * __ldcg(xin + sci[kk]);  // L2: written by other CTAs
        const double dv = c1 * sd[i] + c2 * sinv[i] * (sb[i] - s);
        const double xo = sx[i] + dv;
        sd[i] = dv;
        sx[i] = xo;
        xout[r0 + i] = xo;
      }
      p ^= 1;
    }
    // the next repeat overwrites buffer 0 while slower CTAs may still gather from it
    if (rep + 1 < repeat) grid_barrier(bar, ++gen * gridDim.x);
  }
  for (int i = tid; i < nr; i += kBlock) y[r0 + i] = a * sx[i] + (add ? add[r0 + i] : 0.0);
}

// y = a x + b y   (a, b host scalars; x may alias y only if a-term unused)
__global__ void __launch_bounds__(kBlock) k_axpby(long long n, double a, const double *__restrict__ x, double b,
                                                   double *__restrict__ y) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    y[i] = a * x[i] + (b == 0.0 ? 0.0 : b * y[i]);
}
// y = a * d .* x  (DiagonalMatrix::vmult with a scale, K4)
__global__ void __launch_bounds__(kBlock) k_diag_scale(long long n, double a, const double *__restrict__ d,
                                                        const double *__restrict__ x, double *__restrict__ y) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    y[i] = a * (d ? d[i] * x[i] : x[i]);
}
// y = x / sqrt(*s2) or y = x / *s   (normalisation with a device scalar)
__global__ void __launch_bounds__(kBlock) k_scale_by_inv(long long n, const double *__restrict__ x,
                                                          const double *__restrict__ s, int is_squared,
                                                          double *__restrict__ y) {
  const double nv = is_squared ? sqrt(*s) : *s;
  const double inv = nv != 0.0 ? 1.0 / nv : 0.0;
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock)
    y[i] = x[i] * inv;
}
// first Chebyshev step from a zero guess: d = D^-1 b / theta ; x = d
__global__ void __launch_bounds__(kBlock) k_cheb_zero(long long n, const double *__restrict__ b,
                                                       const double *__restrict__ invd, double inv_theta,
                                                       double *__restrict__ d, double *__restrict__ x) {
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const double dv = invd[i] * b[i] * inv_theta;
    d[i] = dv;
    x[i] = dv;
  }
}

// h[i] = V_i . w for i < nvec, h[nvec] = w . w   (one pass over w per group of
// kMultiDotGroup basis vectors; batched classical Gram-Schmidt of FGMRES, K10)
__global__ void __launch_bounds__(kBlock) k_multidot(long long n, long long n_dot, const double *__restrict__ w,
                                                      const double *__restrict__ V, long long ldv, int nvec,
                                                      int with_norm, Reducer R) {
  __shared__ double smem[32];
  const int n_out = nvec + (with_norm ? 1 : 0);
  for (int g0 = 0; g0 < nvec; g0 += kMultiDotGroup) {
    double acc[kMultiDotGroup];
#pragma unroll
    for (int j = 0; j < kMultiDotGroup; ++j) acc[j] = 0.0;
    const int ng = min(kMultiDotGroup, nvec - g0);
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n_dot; i += (long long)gridDim.x * kBlock) {
      const double wi = w[i];
#pragma unroll
      for (int j = 0; j < kMultiDotGroup; ++j)
        if (j < ng) acc[j] += wi * V[(size_t)(g0 + j) * ldv + i];
    }
    for (int j = 0; j < ng; ++j) reduce_finalize(acc[j], R, g0 + j, n_out, smem);
  }
  if (with_norm) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n_dot; i += (long long)gridDim.x * kBlock)
      s += w[i] * w[i];
    reduce_finalize(s, R, nvec, n_out, smem);
  }
}
// w += sign * sum_i coef[i] V_i     (coef on the device)
__global__ void __launch_bounds__(kBlock) k_multiaxpy(long long n, double *__restrict__ w, const double *__restrict__ V,
                                                       long long ldv, int nvec, const double *__restrict__ coef,
                                                       double sign) {
  __shared__ double c[128];
  for (int j = threadIdx.x; j < nvec; j += kBlock) c[j] = sign * coef[j];
  __syncthreads();
  for (long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    double s = w[i];
    for (int j = 0; j < nvec; ++j) s += c[j] * V[(size_t)j * ldv + i];
    w[i] = s;
  }
}
// h += h2 (tiny)
__global__ void k_small_add(int n, double *a, const double *b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// ---- dense Gauss-Jordan inversion of the coarsest operator (setup, once) ----------
// aug is [A | I] row-major with ld = 2n
__global__ void k_gj_pivot(int n, int k, const double *aug, int *piv_row, double *piv_val, int *singular) {
  __shared__ double sval[kBlock];
  __shared__ int sidx[kBlock];
  double best = -1.0;
  int bi = k;
  for (int r = k + threadIdx.x; r < n; r += kBlock) {
    const double a = fabs(aug[(size_t)r * 2 * n + k]);
    if (a > best) {
      best = a;
      bi = r;
    }
  }
  sval[threadIdx.x] = best;
  sidx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = kBlock / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      if (sval[threadIdx.x + o] > sval[threadIdx.x] ||
          (sval[threadIdx.x + o] == sval[threadIdx.x] && sidx[threadIdx.x + o] < sidx[threadIdx.x])) {
        sval[threadIdx.x] = sval[threadIdx.x + o];
        sidx[threadIdx.x] = sidx[threadIdx.x + o];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *piv_row = sidx[0];
    *piv_val = aug[(size_t)sidx[0] * 2 * n + k];
    if (!(sval[0] > 0.0)) *singular = 1;
  }
}
__global__ void k_gj_swap_scale(int n, int k, double *aug, const int *piv_row, const double *piv_val) {
  const int p = *piv_row;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double *rk = aug + (size_t)k * 2 * n, *rp = aug + (size_t)p * 2 * n;
  const double piv = *piv_val;
  if (c < 2 * n) {
    const double a = rp[c], b = rk[c];
    if (p != k) rp[c] = b;
    rk[c] = a / piv;
  }
}
__global__ void k_gj_save_col(int n, int k, const double *aug, double *col) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) col[r] = (r == k) ? 0.0 : aug[(size_t)r * 2 * n + k];
}
__global__ void k_gj_eliminate(int n, int k, double *aug, const double *col) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c >= 2 * n) return;
  const double f = col[r];
  if (f == 0.0) return;
  aug[(size_t)r * 2 * n + c] -= f * aug[(size_t)k * 2 * n + c];
}
__global__ void k_gj_extract(int n, const double *aug, double *inv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (c < n) inv[(size_t)r * n + c] = aug[(size_t)r * 2 * n + n + c];
}
__global__ void k_dense_from_csr(int n, const int *rp, const int *ci, const double *v, double *aug) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  for (int k = rp[r]; k < rp[r + 1]; ++k) aug[(size_t)r * 2 * n + ci[k]] += v[k];
  aug[(size_t)r * 2 * n + n + r] = 1.0;
}
__global__ void k_inv_diag_from_csr(int n, const int *rp, const int *ci, const double *v, double *invd) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double d = 0.0;
  for (int k = rp[r]; k < rp[r + 1]; ++k)
    if (ci[k] == r) d += v[k];
  invd[r] = 1.0 / d;
}
__global__ void k_fill(long long n, double *x, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = v;
}
__global__ void k_reciprocal(long long n, double *x) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = 1.0 / x[i];
}

}  // namespace fdal
