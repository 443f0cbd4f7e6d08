// fdal.cu — context, device-resident solve drivers and the C ABI (include/fdal.h).
//
// Layout in HBM: every matrix is CSR with int32 row_ptr/col and FP64 values in
// three plain arrays (entries kept in the caller's order); the block system vector
// is ONE contiguous array [block0|block1|block2]; the FGMRES bases V (restart+1)
// and Z (restart) are column-major N x k slabs; AMG levels own x/b/r/d scratch.
// All per-iteration work is enqueued on the context's stream; the host only reads
// back one scalar per inner CG iteration (the residual norm deal.II's SolverControl
// tests) and the Hessenberg column per outer iteration.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen()ed at fdal_comm_init, never linked

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <omp.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/fdal.h"
#include "bsr_build.h"
#include "host_finalize.h"
#include "kernels.cuh"

namespace fdal {

// ------------------------------------------------------------------ NCCL, loaded lazily
// The single-GPU path has no NCCL dependency.  With several ranks the process usually
// already holds torch's bundled libnccl.so.2; dlopen() by soname reuses that copy.
struct NcclApi {
  void *handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};
static NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return &api;
  tried = true;
  api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) api.handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.handle) return &api;
#define FDAL_SYM(name) api.name = (decltype(api.name))dlsym(api.handle, "nccl" #name)
  FDAL_SYM(GetUniqueId);
  FDAL_SYM(CommInitRank);
  FDAL_SYM(CommDestroy);
  FDAL_SYM(AllReduce);
  FDAL_SYM(AllGather);
  FDAL_SYM(Send);
  FDAL_SYM(Recv);
  FDAL_SYM(GroupStart);
  FDAL_SYM(GroupEnd);
  FDAL_SYM(GetErrorString);
#undef FDAL_SYM
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Send && api.Recv &&
           api.GroupStart && api.GroupEnd;
  return &api;
}

// one exchange channel (kernels.cuh: ChanDev) as the host sees it
struct Chan {
  ChanDev d;                   // device view, passed by value to the push / collect kernels
  ChanDev *d_dev = nullptr;    // the same view in device memory (Reducer.ar, XVec.ch)
  bool bcast = false;          // false: halo gather channel; true: all-gather / all-reduce channel
  int cap = 0;                 // doubles per slot
  std::vector<int> nb;         // neighbour ranks (bcast: all ranks, self included)
  std::vector<int> nb_begin;   // gather: my send-list range per neighbour (nnb + 1)
  std::vector<int64_t> src_off;  // [nranks] where rank q's entries start inside MY slot (-1: none)
  size_t recv_off = 0, flags_off = 0, misc_off = 0;  // byte offsets inside the arena
};

struct DevCsr {
  CsrDev d;
  int *rp = nullptr, *ci = nullptr;
  double *v = nullptr;
  bool set = false;
  // BSR copy (node-interleaved vector blocks): replaces the scalar arrays in the SpMV kernels
  BsrDev bsr;
  int bsr_b = 1;
  bool use_bsr = false;
  long long bsr_nblocks = 0;
  // multi-GPU
  bool dist_rows = false;  // rows are a partition: fused reductions need an all-reduce
  bool has_plan = false;
  int n_owned = std::numeric_limits<int>::max(), n_halo = 0, n_send = 0;
  double *halo_buf = nullptr, *send_buf = nullptr;  // NCCL fallback only
  int *send_idx = nullptr;
  std::vector<int> send_counts, recv_counts;
  // merge-path variant (k_spmv_merge): chosen at upload for matrices the row-group kernel cannot spread
  // over the GPU (few rows, or very uneven rows); linear epilogues only
  bool use_merge = false;
  int merge_tiles = 0;
  int *carry_row = nullptr;
  double *carry_val = nullptr;
  int chan = -1;            // index into fdal_ctx::chans (peer-channel mode)
  int *order = nullptr;     // chunk order: interior chunks first (peer-channel mode)
  int n_interior = 0, n_chunks = 0;
};

struct AmgLevel {
  HostCsr hA, hP, hR;
  std::vector<double> h_invd;
  DevCsr A, P, R;
  double *invd = nullptr;
  double lmax = 1.0, ratio = 10.0;
  int degree = 2;
  int n = 0;
  double *xa = nullptr, *xb = nullptr, *b = nullptr, *r = nullptr, *d = nullptr;
};
struct Amg {
  std::vector<AmgLevel> lev;
  double *cinv = nullptr;   // dense inverse of the coarsest operator
  int cn_global = 0;        // rows of the coarsest operator
  // multi-GPU: levels [0, rep_from) are row-partitioned, levels >= rep_from are replicated on
  // every rank (agglomeration: no halos, no collectives down there).  The restricted residual
  // of level rep_from - 1 is all-gathered; rows [c_lo, c_hi) of level rep_from are this rank's.
  int rep_from = -1;
  int64_t c_lo = 0, c_hi = -1;
  double *rown = nullptr;   // owned rows of the restricted residual before the all-gather
  int gather_chan = -1;
  bool dist = false;
  bool ready = false;
};

struct ControlState {
  fdal_control c;
  double initial = 0, reduced_tol = 0;
  int last_step = 0;
  double last_value = 0;
};
enum { ST_ITERATE = 0, ST_SUCCESS = 1, ST_FAILURE = 2 };
static int control_check(ControlState &s, int step, double val) {
  s.last_step = step;
  s.last_value = val;
  if (s.c.type == FDAL_CONTROL_REDUCTION) {
    if (step == 0) {
      s.initial = val;
      s.reduced_tol = val * s.c.reduce;
    }
    if (val < s.reduced_tol) return ST_SUCCESS;
  } else if (s.c.type == FDAL_CONTROL_ITERATION_NUMBER) {
    if (step >= s.c.max_steps) return ST_SUCCESS;
  }
  if (val <= s.c.tol) return ST_SUCCESS;
  if (step >= s.c.max_steps || std::isnan(val)) return ST_FAILURE;
  return ST_ITERATE;
}

struct CgWs {
  int64_t n = 0;
  double *r = nullptr, *z = nullptr, *p = nullptr, *v = nullptr;
  double *x = nullptr;     // the iterate lives here so the iteration body is pointer-stable
  double *bin = nullptr;   // staged right-hand side of the fixed-count mass solve
  double *scal = nullptr;  // S_COUNT device scalars
  bool dist = false;       // vector is partitioned: dots need an all-reduce
  int64_t n_dot = 0;       // entries this rank contributes to dots (replicated tail counted on rank 0)
  // CUDA graphs: one iteration body (host-checked CG) / one whole fixed-count solve
  cudaGraphExec_t body_exec = nullptr, fixed_exec = nullptr;
  int64_t body_nodes = 0, fixed_nodes = 0;
  bool graph_ok = true;
};

// Exact mass solve as a fixed-count Chebyshev iteration on D^-1 M (no inner products).
//   mode 1: one fused SpMV kernel per iteration (k_spmv<.., EpiCheb>), replayed as a CUDA graph
//   mode 2: the whole solve in one persistent kernel with the matrix slice of every CTA staged in
//           shared memory and one grid barrier per iteration (k_mass_cheb_grid)
struct MassCheb {
  int mode = 0;
  int its = 0;
  double lo = 0.0, hi = 0.0;   // spectral interval of D^-1 M used for the coefficients
  double verified_resid = 0.0; // |b - M x| / |b| of the calibration right-hand side
  std::vector<double> coef;    // c1_k, c2_k
  double *d_coef = nullptr, *xbuf = nullptr;
  unsigned int *bar = nullptr;
  int grid = 0, rpc = 0, nnz_cap = 0;
  size_t smem = 0;
  cudaGraphExec_t exec = nullptr;  // mode 1: the whole solve (w.bin -> w.z)
  int64_t exec_nodes = 0;
  bool graph_ok = true;
};

}  // namespace fdal

using namespace fdal;

struct fdal_ctx {
  fdal_config cfg;
  int sms = 148;
  cudaStream_t stream = nullptr;
  // second stream: A x runs beside the single-CTA exact mass solve of the augmented apply
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_mass = true;  // FDAL_OVERLAP=0 switches the fork/join off
  HostCsr hmat[FDAL_MAT_COUNT];
  DevCsr dmat[FDAL_MAT_COUNT];
  std::vector<double> h_winv, h_mp_lumped;
  double *d_winv = nullptr, *d_mp_lumped = nullptr, *d_m_invdiag = nullptr, *d_mp_invdiag = nullptr;
  Amg amg[2];
  bool finalized = false;
  int64_t n0 = 0, n1 = 0, n2 = 0, N = 0, m = 0;
  int nblocks = 0;
  // reductions
  double *d_partials = nullptr;
  unsigned int *d_counter = nullptr;
  double *d_scal = nullptr;   // general scratch scalars (128)
  double *h_scal = nullptr;   // pinned mirror (128)
  // workspaces
  CgWs cg11, cg22, cgblk, cgmass_m, cgmass_p;
  double *t_m0 = nullptr, *t_m1 = nullptr, *t_m2 = nullptr, *t_mw = nullptr;  // m-vectors
  double *t_n0 = nullptr;                                     // n0-vector
  double *t_p0 = nullptr, *t_p1 = nullptr;                    // n1-vectors
  double *t_N0 = nullptr, *t_N1 = nullptr, *t_N2 = nullptr;   // N-vectors (API staging)
  double *V = nullptr, *Z = nullptr, *d_h = nullptr, *d_y = nullptr;
  double *mr_u[3] = {nullptr, nullptr, nullptr}, *mr_m[3] = {nullptr, nullptr, nullptr}, *mr_v = nullptr;
  char *flush_buf = nullptr;
  size_t flush_bytes = 0;
  double *mass_cta_ws = nullptr;       // 5*m scratch of k_mass_pcg_cta (m <= kMassCtaMaxRows)
  double *d_winv_dense = nullptr;      // m x m exact W^-1 (opt-in FDAL_DENSE_WINV=1, m <= kDenseWinvMaxRows)
  int mass_its_m = 0, mass_its_p = 0;  // calibrated fixed iteration counts (exact mass solves)
  MassCheb mcheb_m, mcheb_p;           // Chebyshev form of the same solves (mode 0: keep the Jacobi-PCG)
  // counters
  int its_a11 = 0, its_a22 = 0, its_mass = 0, n_inner_solves = 0;
  int64_t launches = 0, graph_launches = 0;
  // multi-GPU
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;  // setup-time exchanges; per-iteration traffic only in the NCCL fallback mode
  int64_t n_dot_outer = 0;
  // peer channels (default with several ranks): every rank maps every other rank's arena with
  // cudaIpc; halos, scalar / vector all-reduces and the coarse all-gather are direct NVLink
  // stores + flags issued from this library's own kernels (FDAL_COMM=nccl: NCCL fallback)
  bool p2p = false;
  std::vector<Chan> chans;
  char *arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<char *> peer_arena;
  int ch_scalar = -1, ch_vec = -1;
  int vec_cap = 0;
  int spmv_unroll = 4;
  int bsr_unroll2 = 4, bsr_unroll3 = 1;  // blocks per lane in flight for 2x2 / 3x3 blocks (measured, profiles/)
  int fail = 0;
  int bsr_built_on_device = 0, bsr_built_on_host = 0;  // CSR -> BSR conversions at fdal_finalize, by where they ran
  std::vector<void *> allocs;
  std::string err;
};

namespace fdal {

static void set_err(fdal_ctx *c, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  c->err = buf;
}
#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      set_err(c, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
      return FDAL_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

template <class T>
static int dmalloc(fdal_ctx *c, T **p, size_t count) {
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess) {
    set_err(c, "cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
    return FDAL_ERR_ALLOC;
  }
  c->allocs.push_back(q);
  *p = (T *)q;
  return FDAL_OK;
}
static int dvec(fdal_ctx *c, double **p, int64_t n) {
  int st = dmalloc(c, p, (size_t)std::max<int64_t>(n, 1));
  if (st) return st;
  CU(cudaMemsetAsync(*p, 0, (size_t)std::max<int64_t>(n, 1) * sizeof(double), c->stream));
  return FDAL_OK;
}

static int choose_tpr(double avg) {
  const char *env = getenv("FDAL_TPR");
  if (env) {
    int t = atoi(env);
    if (t == 2 || t == 4 || t == 8 || t == 16 || t == 32) return t;
  }
  // ~4 non-zeros per lane (measured on B200: more rows per warp = more independent
  // load chains; see profiles/)
  if (avg <= 12.0) return 2;
  if (avg <= 24.0) return 4;
  if (avg <= 96.0) return 8;
  return 16;
}

// CSR -> BSR with B x B blocks (block-row SoA value layout, see k_bsr_spmv).  Returns false
// (and leaves d untouched) when blocking would store too many explicit zeros.
// FDAL_VERBOSE_SETUP: wall-clock of the phases of fdal_finalize on stderr
struct PhaseTimer {
  bool on = getenv("FDAL_VERBOSE_SETUP") != nullptr;
  double t0 = omp_get_wtime();
  void mark(const char *what, long long items = -1) {
    if (!on) return;
    const double t = omp_get_wtime();
    if (items >= 0)
      fprintf(stderr, "[fdal_finalize] %-44s %8.3f s  (%lld entries)\n", what, t - t0, items);
    else
      fprintf(stderr, "[fdal_finalize] %-44s %8.3f s\n", what, t - t0);
    t0 = t;
  }
};
static int build_bsr(fdal_ctx *c, const HostCsr &h, int b, DevCsr &d, bool *done) {
  *done = false;
  if (b < 2 || b > 3 || h.nr % b || h.nc % b || h.owned_cols() % b || h.nr == 0) return FDAL_OK;
  const int64_t nbr = h.nr / b;
  int64_t nblk = 0;
  int *drp = nullptr, *dcj = nullptr;
  double *dv = nullptr;
  int st;
  bool on_device = false;
  PhaseTimer pt;
  // the scalar CSR arrays are on the device already (upload_csr): convert them there (csrc/bsr_build.cu).
  // FDAL_HOST_BSR=1, or the device routine declining / failing, keeps the OpenMP conversion + a second upload.
  const char *hb = getenv("FDAL_HOST_BSR");
  const bool host_bsr = hb && atoi(hb) > 0;
  if (!host_bsr && d.rp && d.ci && d.v) {
    int max_entries = 0;
#pragma omp parallel for schedule(static) reduction(max : max_entries)
    for (int64_t I = 0; I < nbr; ++I)
      max_entries = std::max(max_entries, h.rp[(size_t)((I + 1) * b)] - h.rp[(size_t)(I * b)]);
    BsrBuilt built;
    const int r = bsr_from_csr_device(c->stream, c->sms, (int)h.nr, h.nnz, d.rp, d.ci, d.v, b, max_entries, 1.35, &built);
    if (r == BSR_BUILD_OK) {
      drp = built.brp;
      dcj = built.bcj;
      dv = built.bv;
      nblk = built.nblk;
      c->allocs.push_back(drp);
      c->allocs.push_back(dcj);
      c->allocs.push_back(dv);
      on_device = true;
    } else if (r == BSR_BUILD_DECLINED) {
      // zero fill too high / sizes: the host routine applies the same rules; fall through and let it decide
    }
  }
  if (!on_device) {
    IntBuf brp, bcj;
    DblBuf bv;
    if (!host_bsr_convert(h, b, 1.35, brp, bcj, bv)) return FDAL_OK;
    nblk = brp[(size_t)nbr];
    if ((st = dmalloc(c, &drp, (size_t)nbr + 1))) return st;
    if ((st = dmalloc(c, &dcj, (size_t)nblk))) return st;
    if ((st = dmalloc(c, &dv, (size_t)nblk * b * b))) return st;
    CU(cudaMemcpyAsync(drp, brp.data(), ((size_t)nbr + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dcj, bcj.data(), (size_t)nblk * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(dv, bv.data(), (size_t)nblk * b * b * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  pt.mark(on_device ? "  CSR -> BSR on the device" : "  CSR -> BSR on the host + upload", h.nnz);
  c->bsr_built_on_device += on_device ? 1 : 0;
  c->bsr_built_on_host += on_device ? 0 : 1;
  d.bsr.nbrows = (int)nbr;
  d.bsr.rp = drp;
  d.bsr.cj = dcj;
  d.bsr.v = dv;
  const double avg = (double)nblk / (double)nbr;
  // measured on B200 (profiles/r2_kernel_probe.md): 2x2 blocks, 9-25 per row: 4 lanes x 4 blocks in
  // flight; 3x3 blocks, ~64 per row: 16 lanes, one block each
  d.bsr.tpr = avg <= 6 ? 2 : avg <= 24 ? 4 : 16;
  if (const char *e = getenv("FDAL_BSR_TPR")) {
    const int t = atoi(e);
    if (t == 2 || t == 4 || t == 8 || t == 16) d.bsr.tpr = t;
  }
  if (d.bsr.tpr < b) d.bsr.tpr = 4;  // the B epilogue lanes of a row group must exist
  d.bsr_b = b;
  d.bsr_nblocks = nblk;
  d.use_bsr = true;
  *done = true;
  return FDAL_OK;
}

// chunk order of a row-partitioned matrix: chunks (the rows one CTA handles per loop trip)
// whose rows only touch owned columns come first; the SpMV kernels acquire the halo before
// the first chunk at a position >= n_interior, so the peers' pushes overlap the interior rows
static int build_chunk_order(fdal_ctx *c, const HostCsr &h, DevCsr &d) {
  const int unit = d.use_bsr ? d.bsr_b : 1;
  const int tpr = d.use_bsr ? d.bsr.tpr : d.d.tpr;
  const int64_t rpb = (int64_t)(kBlock / tpr) * unit;  // scalar rows per chunk
  const int64_t nchunks = (h.nr + rpb - 1) / rpb;
  std::vector<char> bnd((size_t)nchunks, 0);
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < nchunks; ++q) {
    const int64_t r0 = q * rpb, r1 = std::min<int64_t>(h.nr, r0 + rpb);
    char f = 0;
    for (int k = h.rp[r0]; k < h.rp[r1] && !f; ++k) f = h.ci[k] >= h.n_owned;
    bnd[(size_t)q] = f;
  }
  std::vector<int> order;
  order.reserve((size_t)nchunks);
  for (int64_t q = 0; q < nchunks; ++q)
    if (!bnd[(size_t)q]) order.push_back((int)q);
  d.n_interior = (int)order.size();
  d.n_chunks = (int)nchunks;
  for (int64_t q = 0; q < nchunks; ++q)
    if (bnd[(size_t)q]) order.push_back((int)q);
  int st;
  if ((st = dmalloc(c, &d.order, (size_t)nchunks))) return st;
  if (nchunks) CU(cudaMemcpyAsync(d.order, order.data(), (size_t)nchunks * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return FDAL_OK;
}

static int upload_csr(fdal_ctx *c, const HostCsr &h, DevCsr &d, int bsr_b = 1) {
  if (h.nnz >= (int64_t)std::numeric_limits<int>::max() || h.nr >= (int64_t)std::numeric_limits<int>::max()) {
    set_err(c, "matrix with %lld nnz exceeds the 32-bit row_ptr of this build", (long long)h.nnz);
    return FDAL_ERR_UNSUPPORTED;
  }
  int st;
  if ((st = dmalloc(c, &d.rp, (size_t)h.nr + 1))) return st;
  if ((st = dmalloc(c, &d.ci, (size_t)h.nnz))) return st;
  if ((st = dmalloc(c, &d.v, (size_t)h.nnz))) return st;
  CU(cudaMemcpyAsync(d.rp, h.rp.data(), ((size_t)h.nr + 1) * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (h.nnz) {
    CU(cudaMemcpyAsync(d.ci, h.ci.data(), (size_t)h.nnz * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d.v, h.v.data(), (size_t)h.nnz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  CU(cudaStreamSynchronize(c->stream));
  d.d.nrows = (int)h.nr;
  d.d.ncols = (int)h.nc;
  d.d.nnz = h.nnz;
  d.d.rp = d.rp;
  d.d.ci = d.ci;
  d.d.v = d.v;
  d.d.tpr = choose_tpr(h.nr ? (double)h.nnz / (double)h.nr : 1.0);
  d.set = true;
  if (bsr_b > 1 && !getenv("FDAL_NO_BSR")) {
    bool done = false;
    if ((st = build_bsr(c, h, bsr_b, d, &done))) return st;
  }
  if (!d.use_bsr && !h.has_plan && h.nr > 0 && h.nnz > 0) {
    // merge-path variant: when the row groups of k_spmv would fill less than half of the GPU's thread
    // slots, or when one row is far longer than the average (FDAL_MERGE=1 / 0: always / never)
    int maxlen = 0;
    for (int64_t i = 0; i < h.nr; ++i) maxlen = std::max(maxlen, h.rp[i + 1] - h.rp[i]);
    const double avg = (double)h.nnz / (double)h.nr;
    bool want = (h.nnz >= 200000 && (double)h.nr * d.d.tpr < 0.5 * c->sms * 2048.0) || (h.nnz >= 1000000 && maxlen > 64.0 * avg);
    if (const char *e = getenv("FDAL_MERGE")) want = atoi(e) > 0;
    if (want) {
      d.merge_tiles = (int)((h.nr + h.nnz + kMergeTile - 1) / kMergeTile);
      if ((st = dmalloc(c, &d.carry_row, (size_t)d.merge_tiles))) return st;
      if ((st = dmalloc(c, &d.carry_val, (size_t)d.merge_tiles))) return st;
      d.use_merge = true;
    }
  }
  if (h.has_plan) {
    d.has_plan = true;
    d.n_owned = (int)h.n_owned;
    d.n_halo = (int)h.n_halo;
    d.n_send = (int)h.send_idx.size();
    d.send_counts = h.send_counts;
    d.recv_counts = h.recv_counts;
    if ((st = dmalloc(c, &d.send_idx, (size_t)d.n_send))) return st;
    if (d.n_send)
      CU(cudaMemcpyAsync(d.send_idx, h.send_idx.data(), (size_t)d.n_send * sizeof(int), cudaMemcpyHostToDevice,
                         c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->p2p) {
      // register a gather channel: neighbours = ranks I send to or receive from (a symmetric relation)
      Chan ch;
      ch.bcast = false;
      ch.cap = d.n_halo;
      ch.src_off.assign((size_t)c->nranks, -1);
      ch.nb_begin.push_back(0);
      int64_t ro = 0;
      int so = 0;
      for (int q = 0; q < c->nranks; ++q) {
        if (h.recv_counts[q]) ch.src_off[(size_t)q] = ro;
        ro += h.recv_counts[q];
        if (h.send_counts[q] || h.recv_counts[q]) {
          ch.nb.push_back(q);
          so += h.send_counts[q];
          ch.nb_begin.push_back(so);
        }
      }
      d.chan = (int)c->chans.size();
      c->chans.push_back(ch);
      if ((st = build_chunk_order(c, h, d))) return st;
    } else {
      if ((st = dvec(c, &d.halo_buf, d.n_halo))) return st;
      if ((st = dvec(c, &d.send_buf, d.n_send))) return st;
    }
  }
  return FDAL_OK;
}

// ------------------------------------------------------------------ launch helpers
static inline int grid_rows(const fdal_ctx *c, long long nrows, int tpr) {
  const long long rpb = kBlock / tpr;
  long long g = (nrows + rpb - 1) / rpb;
  return (int)std::max<long long>(1, std::min<long long>(g, (long long)c->sms * kMaxGridPerSM));
}
static inline int grid_elems(const fdal_ctx *c, long long n) {
  long long g = (n + kBlock - 1) / kBlock;
  return (int)std::max<long long>(1, std::min<long long>(g, (long long)c->sms * kMaxGridPerSM));
}
// dist: the value is a partial over this rank's rows -> the finishing block of the kernel
// all-reduces it over the ranks through the scalar channel (peer-channel mode)
static inline Reducer reducer(fdal_ctx *c, double *out, bool dist = false) {
  Reducer R{c->d_partials, c->d_counter, out, nullptr};
  if (dist && c->p2p) R.ar = c->chans[(size_t)c->ch_scalar].d_dev;
  return R;
}

static void launch_check(fdal_ctx *c) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess && !c->fail) {
    set_err(c, "kernel launch failed: %s", cudaGetErrorString(e));
    c->fail = FDAL_ERR_CUDA;
  }
}
// NCCL fallback: sum of device doubles over all ranks, in stream order
static void nccl_allreduce(fdal_ctx *c, double *p, size_t count) {
  if (c->nranks <= 1 || !count) return;
  ncclResult_t r = nccl_api()->AllReduce(p, p, count, ncclDouble, ncclSum, c->comm, c->stream);
  if (r != ncclSuccess && !c->fail) c->fail = FDAL_ERR_NCCL;
}
// after a kernel with a fused reduction over partitioned rows: in peer-channel mode the kernel
// has already all-reduced the scalar itself
static void finish_scalar(fdal_ctx *c, double *p, size_t count, bool dist) {
  if (dist && c->nranks > 1 && !c->p2p) nccl_allreduce(c, p, count);
}
static int chan_grid(fdal_ctx *c, int n) { return std::max(1, std::min((n + kBlock - 1) / kBlock, c->sms)); }
// in-place sum of a replicated-size vector (C x partials, count <= vec_cap) over all ranks
static void allreduce_vec(fdal_ctx *c, double *p, size_t count) {
  if (c->nranks <= 1 || !count) return;
  if (!c->p2p) {
    nccl_allreduce(c, p, count);
    return;
  }
  const Chan &ch = c->chans[(size_t)c->ch_vec];
  k_chan_push_bcast<<<dim3(chan_grid(c, (int)count), c->nranks), kBlock, 0, c->stream>>>(ch.d, p, (int)count);
  k_chan_collect<true><<<chan_grid(c, (int)count), kBlock, 0, c->stream>>>(ch.d, p, (int)count, c->vec_cap);
  c->launches += 2;
}
// all-gather of the owned rows of a vector into the replicated vector `full` (n_full entries)
static void allgather_rows(fdal_ctx *c, Amg &g, const double *own, int n_own, double *full, int n_full) {
  if (!c->p2p) {
    cudaMemsetAsync(full, 0, (size_t)n_full * sizeof(double), c->stream);
    if (n_own)
      cudaMemcpyAsync(full + g.c_lo, own, (size_t)n_own * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
    nccl_allreduce(c, full, (size_t)n_full);
    return;
  }
  const Chan &ch = c->chans[(size_t)g.gather_chan];
  k_chan_push_bcast<<<dim3(chan_grid(c, n_own), c->nranks), kBlock, 0, c->stream>>>(ch.d, own, n_own);
  k_chan_collect<false><<<chan_grid(c, n_full), kBlock, 0, c->stream>>>(ch.d, full, n_full, 0);
  c->launches += 2;
}
// make the owners' entries of x visible in A's halo: peer-channel mode pushes my entries straight
// into the neighbours' receive slots (one kernel; the SpMV that follows acquires the flags before its
// first boundary chunk); NCCL fallback: pack -> grouped send/recv
static void halo_exchange(fdal_ctx *c, const DevCsr &A, const double *x) {
  if (!A.has_plan || c->nranks <= 1) return;
  if (c->p2p) {
    const Chan &ch = c->chans[(size_t)A.chan];
    if (ch.nb.empty()) return;
    k_chan_push_gather<<<chan_grid(c, A.n_send), kBlock, 0, c->stream>>>(ch.d, A.send_idx, x);
    c->launches++;
    return;
  }
  if (A.n_send) {
    k_pack<<<grid_elems(c, A.n_send), kBlock, 0, c->stream>>>(A.n_send, A.send_idx, x, A.send_buf);
    c->launches++;
  }
  NcclApi *n = nccl_api();
  n->GroupStart();
  size_t so = 0, ro = 0;
  for (int q = 0; q < c->nranks; ++q) {
    if (A.send_counts[q]) n->Send(A.send_buf + so, (size_t)A.send_counts[q], ncclDouble, q, c->comm, c->stream);
    if (A.recv_counts[q]) n->Recv(A.halo_buf + ro, (size_t)A.recv_counts[q], ncclDouble, q, c->comm, c->stream);
    so += (size_t)A.send_counts[q];
    ro += (size_t)A.recv_counts[q];
  }
  ncclResult_t r = n->GroupEnd();
  if (r != ncclSuccess && !c->fail) c->fail = FDAL_ERR_NCCL;
}
// `grid`: CTAs of the launch.  In peer-channel mode the first g_bnd CTAs acquire the halo and share the
// boundary chunks, the rest share the interior chunks (kernels.cuh: chunk_range)
static inline XVec xv(const fdal_ctx *c, const DevCsr &A, const double *x, int *grid) {
  XVec X{x, A.halo_buf, A.n_owned};
  if (c->p2p && A.chan >= 0 && !c->chans[(size_t)A.chan].nb.empty()) {
    X.ch = c->chans[(size_t)A.chan].d_dev;
    X.order = A.order;
    X.n_interior = A.n_interior;
    const int n_bnd = A.n_chunks - A.n_interior;
    const int g = std::max(1, std::min(*grid, A.n_chunks));
    // Default: every CTA acquires the halo before its loop and all CTAs share all chunks.  The interior /
    // boundary CTA split (FDAL_SPLIT=1: boundary CTAs acquire, interior CTAs start at once) was measured
    // and lost: 3-D nel=40 on 2 B200s, V-cycle 2382 us with the split, 1896 us without (profiles/r2_multi_gpu.md)
    // — the ranks run in lockstep, the flags are there when the kernel starts, and two CTA populations
    // with different work per chunk finish at different times.
    static const bool split = getenv("FDAL_SPLIT") != nullptr && atoi(getenv("FDAL_SPLIT")) > 0;
    if (!split) {
      X.n_interior = 0;
      X.g_bnd = g;
    } else if (n_bnd == 0)
      X.g_bnd = 0;
    else if (A.n_interior == 0 || g < 2)
      X.g_bnd = g;
    else  // CTAs in proportion to the chunks of each kind, at least one each
      X.g_bnd = std::max(1, std::min(g - 1, (int)((double)g * n_bnd / A.n_chunks + 0.5)));
    *grid = g;
  }
  return X;
}

template <class Epi, bool TWO>
static void spmv_bsr(fdal_ctx *c, const DevCsr &A, const double *x, const DevCsr *B2, const double *t2, Epi epi,
                     double *red_out) {
  const int tpr = A.bsr.tpr;
  int g = grid_rows(c, A.bsr.nbrows, tpr);
  Reducer R = reducer(c, red_out, A.dist_rows);
  XVec X = xv(c, A, x, &g);
  CsrDev b2 = B2 ? B2->d : CsrDev();
  const int unroll = A.bsr_b == 2 ? c->bsr_unroll2 : c->bsr_unroll3;
#define FDAL_BSR_LAUNCH(BB, TT)                                                                      \
  do {                                                                                               \
    if (X.ch) {                                                                                      \
      if (unroll >= 4)                                                                               \
        k_bsr_spmv<BB, TT, Epi, TWO, 4, true><<<g, kBlock, 0, c->stream>>>(A.bsr, X, b2, t2, epi, R); \
      else                                                                                           \
        k_bsr_spmv<BB, TT, Epi, TWO, 1, true><<<g, kBlock, 0, c->stream>>>(A.bsr, X, b2, t2, epi, R); \
    } else if (unroll >= 4)                                                                          \
      k_bsr_spmv<BB, TT, Epi, TWO, 4><<<g, kBlock, 0, c->stream>>>(A.bsr, X, b2, t2, epi, R);        \
    else                                                                                             \
      k_bsr_spmv<BB, TT, Epi, TWO, 1><<<g, kBlock, 0, c->stream>>>(A.bsr, X, b2, t2, epi, R);        \
  } while (0)
  if (A.bsr_b == 2) {
    switch (tpr) {
      case 2: FDAL_BSR_LAUNCH(2, 2); break;
      case 4: FDAL_BSR_LAUNCH(2, 4); break;
      case 16: FDAL_BSR_LAUNCH(2, 16); break;
      default: FDAL_BSR_LAUNCH(2, 8); break;
    }
  } else {
    switch (tpr) {
      case 4: FDAL_BSR_LAUNCH(3, 4); break;
      case 16: FDAL_BSR_LAUNCH(3, 16); break;
      default: FDAL_BSR_LAUNCH(3, 8); break;
    }
  }
#undef FDAL_BSR_LAUNCH
  c->launches++;
  launch_check(c);
  finish_scalar(c, red_out, red_out ? 1 : 0, A.dist_rows);
}
template <class Epi>
static void spmv(fdal_ctx *c, const DevCsr &A, const double *x, Epi epi, double *red_out = nullptr) {
  halo_exchange(c, A, x);
  if (A.d.nrows == 0) {
    if (red_out) {
      if (A.dist_rows && c->p2p) {  // still take part in the fused all-reduce
        k_dot<<<1, kBlock, 0, c->stream>>>(0, x, x, reducer(c, red_out, true));
        c->launches++;
      } else {
        cudaMemsetAsync(red_out, 0, sizeof(double), c->stream);
        finish_scalar(c, red_out, 1, A.dist_rows);
      }
    }
    return;
  }
  if (A.use_bsr) {
    spmv_bsr<Epi, false>(c, A, x, nullptr, nullptr, epi, red_out);
    return;
  }
  if constexpr (std::is_same<Epi, EpiAssign>::value || std::is_same<Epi, EpiAdd>::value) {
    if (A.use_merge && !red_out) {
      k_spmv_merge<<<A.merge_tiles, kBlock, 0, c->stream>>>(A.d, x, epi.alpha, std::is_same<Epi, EpiAdd>::value ? 1 : 0,
                                                            epi.y, A.carry_row, A.carry_val);
      k_spmv_merge_fixup<<<std::max(1, std::min((A.merge_tiles + kBlock - 1) / kBlock, c->sms)), kBlock, 0, c->stream>>>(
          A.merge_tiles, A.d.nrows, epi.alpha, A.carry_row, A.carry_val, epi.y);
      c->launches += 2;
      launch_check(c);
      return;
    }
  }
  int g = grid_rows(c, A.d.nrows, A.d.tpr);
  Reducer R = reducer(c, red_out, A.dist_rows);
  XVec X = xv(c, A, x, &g);
#define FDAL_CSR_LAUNCH(TT)                                                         \
  do {                                                                              \
    if (X.ch)                                                                       \
      k_spmv<TT, Epi, 4, true><<<g, kBlock, 0, c->stream>>>(A.d, X, epi, R);        \
    else if (c->spmv_unroll > 1)                                                    \
      k_spmv<TT, Epi, 4><<<g, kBlock, 0, c->stream>>>(A.d, X, epi, R);              \
    else                                                                            \
      k_spmv<TT, Epi><<<g, kBlock, 0, c->stream>>>(A.d, X, epi, R);                 \
  } while (0)
  switch (A.d.tpr) {
    case 2: FDAL_CSR_LAUNCH(2); break;
    case 4: FDAL_CSR_LAUNCH(4); break;
    case 8: FDAL_CSR_LAUNCH(8); break;
    case 16: FDAL_CSR_LAUNCH(16); break;
    default: FDAL_CSR_LAUNCH(32); break;
  }
#undef FDAL_CSR_LAUNCH
  c->launches++;
  launch_check(c);
  finish_scalar(c, red_out, red_out ? 1 : 0, A.dist_rows);
}
template <class Epi>
static void spmv2(fdal_ctx *c, const DevCsr &A, const double *x, const DevCsr &Ct, const double *t, Epi epi,
                  double *red_out = nullptr) {
  halo_exchange(c, A, x);
  if (A.d.nrows == 0) {
    if (red_out) {
      if (A.dist_rows && c->p2p) {
        k_dot<<<1, kBlock, 0, c->stream>>>(0, x, x, reducer(c, red_out, true));
        c->launches++;
      } else {
        cudaMemsetAsync(red_out, 0, sizeof(double), c->stream);
        finish_scalar(c, red_out, 1, A.dist_rows);
      }
    }
    return;
  }
  if (A.use_bsr) {
    spmv_bsr<Epi, true>(c, A, x, &Ct, t, epi, red_out);
    return;
  }
  int g = grid_rows(c, A.d.nrows, A.d.tpr);
  Reducer R = reducer(c, red_out, A.dist_rows);
  XVec X = xv(c, A, x, &g);
#define FDAL_CSR2_LAUNCH(TT)                                                               \
  do {                                                                                     \
    if (X.ch)                                                                              \
      k_spmv2<TT, Epi, 4, true><<<g, kBlock, 0, c->stream>>>(A.d, X, Ct.d, t, epi, R);     \
    else if (c->spmv_unroll > 1)                                                           \
      k_spmv2<TT, Epi, 4><<<g, kBlock, 0, c->stream>>>(A.d, X, Ct.d, t, epi, R);           \
    else                                                                                   \
      k_spmv2<TT, Epi><<<g, kBlock, 0, c->stream>>>(A.d, X, Ct.d, t, epi, R);              \
  } while (0)
  switch (A.d.tpr) {
    case 2: FDAL_CSR2_LAUNCH(2); break;
    case 4: FDAL_CSR2_LAUNCH(4); break;
    case 8: FDAL_CSR2_LAUNCH(8); break;
    case 16: FDAL_CSR2_LAUNCH(16); break;
    default: FDAL_CSR2_LAUNCH(32); break;
  }
#undef FDAL_CSR2_LAUNCH
  c->launches++;
  launch_check(c);
  finish_scalar(c, red_out, red_out ? 1 : 0, A.dist_rows);
}
// out = a[0:n) . b[0:n); dist: n is this rank's share, all-reduced over the ranks
static void dot(fdal_ctx *c, int64_t n, const double *a, const double *b, double *out, bool dist = false) {
  k_dot<<<grid_elems(c, std::max<int64_t>(n, 1)), kBlock, 0, c->stream>>>(n, a, b, reducer(c, out, dist));
  c->launches++;
  finish_scalar(c, out, 1, dist);
}
static void axpby(fdal_ctx *c, int64_t n, double a, const double *x, double b, double *y) {
  if (n == 0) return;
  k_axpby<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, a, x, b, y);
  c->launches++;
}
static void dscale(fdal_ctx *c, int64_t n, double a, double *y) {  // y *= a
  if (n == 0 || a == 1.0) return;
  k_axpby<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, a, y, 0.0, y);
  c->launches++;
}
static void diag_scale(fdal_ctx *c, int64_t n, double a, const double *d, const double *x, double *y) {
  if (n == 0) return;
  k_diag_scale<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, a, d, x, y);
  c->launches++;
}
static void dcopy(fdal_ctx *c, int64_t n, const double *x, double *y) {
  if (n && x != y) cudaMemcpyAsync(y, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
}
static void dzero(fdal_ctx *c, int64_t n, double *y) {
  if (n) cudaMemsetAsync(y, 0, (size_t)n * sizeof(double), c->stream);
}
// read device scalars to the host (one sync)
static int read_scalars(fdal_ctx *c, const double *d, int count, double *out) {
  CU(cudaMemcpyAsync(c->h_scal, d, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());  // a failed launch must not pass for a converged scalar
  if (c->fail == FDAL_ERR_CUDA) return c->fail;
  memcpy(out, c->h_scal, (size_t)count * sizeof(double));
  return FDAL_OK;
}

// ------------------------------------------------------------------ AMG V-cycle (K8)
static void cheb_coeffs(const AmgLevel &L, double &theta, double &delta, double &s1) {
  const double beta = 1.1 * L.lmax, alpha = L.lmax / L.ratio;
  delta = 0.5 * (beta - alpha);
  theta = 0.5 * (beta + alpha);
  s1 = theta / delta;
}
// Chebyshev(degree) on level L: zero_guess -> result in *xcur; returns the buffer
// holding the result.  final_out (optional): the last step writes there (and fuses
// dot(b, x) into red_out when given).
static double *cheb(fdal_ctx *c, AmgLevel &L, const double *b, double *xcur, bool zero_guess, double *final_out,
                    double *red_out) {
  double theta, delta, s1;
  cheb_coeffs(L, theta, delta, s1);
  double rho = 1.0 / s1;
  const int n = L.n;
  int k = 0;
  double *cur = xcur;
  auto other = [&](double *p) { return p == L.xa ? L.xb : L.xa; };
  if (zero_guess) {
    double *dst = (L.degree == 1 && final_out) ? final_out : cur;
    k_cheb_zero<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, b, L.invd, 1.0 / theta, L.d, dst);
    c->launches++;
    cur = dst;
    k = 1;
  }
  for (; k < L.degree; ++k) {
    const bool first = (k == 0);
    double c1 = 0.0, c2 = 1.0 / theta;
    if (!first) {
      const double rho1 = 1.0 / (2.0 * s1 - rho);
      c1 = rho1 * rho;
      c2 = 2.0 * rho1 / delta;
      rho = rho1;
    }
    const bool last = (k == L.degree - 1);
    double *dst = (last && final_out) ? final_out : other(cur);
    if (last && red_out) {
      EpiCheb<true> e{b, L.invd, cur, L.d, dst, c1, c2, first ? 1 : 0};
      spmv(c, L.A, cur, e, red_out);
    } else {
      EpiCheb<false> e{b, L.invd, cur, L.d, dst, c1, c2, first ? 1 : 0};
      spmv(c, L.A, cur, e);
    }
    cur = dst;
  }
  return cur;
}
// coarsest level: x = A_L^-1 b (dense inverse, L2-resident).  With several ranks the coarsest
// level is always inside the replicated part of the hierarchy: every rank applies the whole inverse.
static void coarse_solve(fdal_ctx *c, Amg &g, const double *b, double *x) {
  AmgLevel &C = g.lev.back();
  if (C.n > 0) {
    k_gemv<<<(C.n * 32 + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(C.n, g.cn_global, g.cinv, b, x);
    c->launches++;
  }
}
// z = AMG(b); optionally *red_out = b.z
static void vcycle(fdal_ctx *c, Amg &g, const double *b0, double *z, double *red_out) {
  const int nl = (int)g.lev.size();
  if (nl == 1) {
    coarse_solve(c, g, b0, z);
    if (red_out) dot(c, g.lev[0].n, b0, z, red_out, false);
    return;
  }
  std::vector<double *> xres(nl, nullptr);
  for (int l = 0; l < nl - 1; ++l) {
    AmgLevel &L = g.lev[l];
    const double *bl = l == 0 ? b0 : L.b;
    double *x = cheb(c, L, bl, L.xa, true, nullptr, nullptr);
    spmv(c, L.A, x, EpiResid{L.r, bl});
    if (g.dist && l + 1 == g.rep_from) {
      // last partitioned level: restrict onto the owned coarse rows, then all-gather the
      // replicated right-hand side of the agglomerated levels
      spmv(c, L.R, L.r, EpiAssign{g.rown, 1.0});
      allgather_rows(c, g, g.rown, (int)(g.c_hi - g.c_lo), g.lev[l + 1].b, g.lev[l + 1].n);
    } else {
      spmv(c, L.R, L.r, EpiAssign{g.lev[l + 1].b, 1.0});
    }
    xres[l] = x;
  }
  {
    AmgLevel &C = g.lev[nl - 1];
    coarse_solve(c, g, C.b, C.xa);
    xres[nl - 1] = C.xa;
  }
  for (int l = nl - 2; l >= 0; --l) {
    AmgLevel &L = g.lev[l];
    const double *bl = l == 0 ? b0 : L.b;
    spmv(c, L.P, xres[l + 1], EpiAdd{xres[l], 1.0});
    xres[l] = cheb(c, L, bl, xres[l], false, l == 0 ? z : nullptr, l == 0 ? red_out : nullptr);
  }
}

// ------------------------------------------------------------------ device CG (K9)
using OpFn = std::function<void(const double *in, double *out, double *dot_out)>;

static void cg_start(fdal_ctx *c, CgWs &w, const double *b, double *x) {
  dzero(c, w.n, x);
  dcopy(c, w.n, b, w.r);
  dzero(c, w.n, w.p);
  k_set_scalar<<<1, 1, 0, c->stream>>>(w.scal + S_RHO_OLD, std::numeric_limits<double>::infinity());
  c->launches++;
}
// one iteration body, no host interaction (capturable)
static void cg_body(fdal_ctx *c, CgWs &w, const OpFn &op, const OpFn &prec, double *x) {
  const int64_t n = w.n;
  prec(w.r, w.z, w.scal + S_RHO);
  k_cg_update_p<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, w.z, w.p, w.scal);
  c->launches++;
  op(w.p, w.v, w.scal + S_PV);
  k_cg_update_xr<<<grid_elems(c, n), kBlock, 0, c->stream>>>(n, w.dist ? w.n_dot : n, w.p, w.v, x, w.r, w.scal,
                                                               reducer(c, w.scal + S_RR, w.dist));
  c->launches++;
  finish_scalar(c, w.scal + S_RR, 1, w.dist);
}
static bool stream_is_capturing(fdal_ctx *c) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(c->stream, &cs);
  return cs != cudaStreamCaptureStatusNone;
}
// Capture `enqueue` (kernel launches on c->stream only, no host interaction) into an
// executable graph.  On failure the caller keeps launching directly.
static bool capture_graph(fdal_ctx *c, const std::function<void()> &enqueue, cudaGraphExec_t *exec, int64_t *nodes) {
  const int64_t l0 = c->launches;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  enqueue();
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &g);
  *nodes = c->launches - l0;
  c->launches = l0;  // nothing ran yet
  if (e != cudaSuccess || !g) {
    cudaGetLastError();
    return false;
  }
  e = cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *exec = nullptr;
    return false;
  }
  return true;
}
static bool graphs_enabled(const fdal_ctx *c) {
  // NCCL send/recv and all-reduce are captured into the graphs as well (verified on 2 GPUs)
  static const bool no_dist_graphs = getenv("FDAL_NO_DIST_GRAPHS") != nullptr;
  return c->cfg.use_graphs && (c->nranks <= 1 || !no_dist_graphs);
}
static void run_cg_body(fdal_ctx *c, CgWs &w, const OpFn &op, const OpFn &prec, bool capturable) {
  if (graphs_enabled(c) && capturable && w.graph_ok && !stream_is_capturing(c)) {
    if (!w.body_exec)
      w.graph_ok = capture_graph(c, [&]() { cg_body(c, w, op, prec, w.x); }, &w.body_exec, &w.body_nodes);
    if (w.body_exec) {
      if (cudaGraphLaunch(w.body_exec, c->stream) != cudaSuccess && !c->fail) c->fail = FDAL_ERR_CUDA;
      c->launches += w.body_nodes;
      c->graph_launches++;
      return;
    }
  }
  cg_body(c, w, op, prec, w.x);
}
// deal.II SolverCG from a zero initial guess (inverse_operator), host-checked control.
// `capturable`: op and prec never synchronise with the host, so one iteration body is
// replayed as a CUDA graph (pointer-stable: the iterate lives in w.x).
static int cg_solve(fdal_ctx *c, CgWs &w, const OpFn &op, const OpFn &prec, const fdal_control &ctl, const double *b,
                    double *x, int *its_out, int fail_code, bool capturable) {
  ControlState cs;
  cs.c = ctl;
  cg_start(c, w, b, w.x);
  dot(c, w.dist ? w.n_dot : w.n, w.r, w.r, w.scal + S_RR, w.dist);
  double rr;
  int st = read_scalars(c, w.scal + S_RR, 1, &rr);
  if (st) return st;
  int state = control_check(cs, 0, std::sqrt(std::fabs(rr)));
  int it = 0;
  while (state == ST_ITERATE) {
    ++it;
    run_cg_body(c, w, op, prec, capturable);
    st = read_scalars(c, w.scal + S_RR, 1, &rr);
    if (st) return st;
    state = control_check(cs, it, std::sqrt(std::fabs(rr)));
  }
  dcopy(c, w.n, w.x, x);
  *its_out = it;
  return state == ST_SUCCESS ? FDAL_OK : fail_code;
}
// fixed-count Jacobi-PCG on a mass matrix: the device replacement of
// SparseDirectUMFPACK::vmult (K5); no host interaction
static void mass_prec(fdal_ctx *c, const double *invdiag, int64_t n, const double *r, double *z, double *dot_out,
                      bool dist = false) {
  k_diag_prec_dot<<<grid_elems(c, std::max<int64_t>(n, 1)), kBlock, 0, c->stream>>>(n, invdiag, r, z,
                                                                                   reducer(c, dot_out, dist));
  c->launches++;
  finish_scalar(c, dot_out, 1, dist);
}
static void mass_solve_fixed(fdal_ctx *c, CgWs &w, const DevCsr &M, const double *invdiag, int its, const double *b,
                             double *x) {
  OpFn op = [&](const double *in, double *out, double *d) { spmv(c, M, in, EpiDotX{out, in}, d); };
  OpFn pr = [&](const double *r, double *z, double *d) { mass_prec(c, invdiag, w.n, r, z, d, w.dist); };
  auto whole = [&]() {
    cg_start(c, w, w.bin, w.x);
    for (int i = 0; i < its; ++i) cg_body(c, w, op, pr, w.x);
  };
  dcopy(c, w.n, b, w.bin);
  bool done = false;
  if (graphs_enabled(c) && w.graph_ok && !stream_is_capturing(c)) {
    if (!w.fixed_exec) w.graph_ok = capture_graph(c, whole, &w.fixed_exec, &w.fixed_nodes);
    if (w.fixed_exec) {
      if (cudaGraphLaunch(w.fixed_exec, c->stream) != cudaSuccess && !c->fail) c->fail = FDAL_ERR_CUDA;
      c->launches += w.fixed_nodes;
      c->graph_launches++;
      done = true;
    }
  }
  if (!done) whole();
  dcopy(c, w.n, w.x, x);
}
static int mass_calibrate(fdal_ctx *c, CgWs &w, const DevCsr &M, const double *invdiag, int *its_out,
                          CgHistory *lanczos = nullptr) {
  // count the iterations Jacobi-PCG needs to push the recursive residual below
  // 1e-17 |b| on a rough right-hand side; the solve then always runs that many + 3
  const int64_t n = w.n;
  std::vector<double> hb((size_t)n);
  unsigned long long s = 0x9E3779B97F4A7C15ull;
  for (int64_t i = 0; i < n; ++i) {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    hb[(size_t)i] = ((double)(s >> 11) / 9007199254740992.0) * 2.0 - 1.0;
  }
  double *b = w.bin, *x = w.x;
  int st;
  CU(cudaMemcpyAsync(b, hb.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  cg_start(c, w, b, x);
  dot(c, n, w.r, w.r, w.scal + S_RR, w.dist);
  double rr0;
  if ((st = read_scalars(c, w.scal + S_RR, 1, &rr0))) return st;
  OpFn op = [&](const double *in, double *out, double *d) { spmv(c, M, in, EpiDotX{out, in}, d); };
  OpFn pr = [&](const double *r, double *z, double *d) { mass_prec(c, invdiag, n, r, z, d, w.dist); };
  const int cap = c->cfg.exact_mass_max_its > 0 ? c->cfg.exact_mass_max_its : 300;
  int it = 0;
  double rr = rr0;
  std::vector<double> rho_k, pv_k;  // CG coefficients -> Lanczos matrix of D^-1 M (bounds for the Chebyshev form)
  while (it < cap && std::sqrt(std::fabs(rr)) > 1e-17 * std::sqrt(rr0)) {
    cg_body(c, w, op, pr, x);
    ++it;
    double sc[4];
    if ((st = read_scalars(c, w.scal, 4, sc))) return st;
    rr = sc[S_RR];
    rho_k.push_back(sc[S_RHO]);
    pv_k.push_back(sc[S_PV]);
  }
  *its_out = std::min(cap, it + 3);
  if (lanczos) {
    lanczos->rho = rho_k;
    lanczos->pv = pv_k;
  }
  if (it >= cap && std::sqrt(std::fabs(rr)) > 1e-17 * std::sqrt(rr0))
    set_err(c, "warning: the exact mass solve stopped at its cap of %d Jacobi-PCG iterations with relative residual %.2e",
            cap, std::sqrt(std::fabs(rr) / rr0));
  return FDAL_OK;
}

// ---- Chebyshev form of the exact mass solves (bounds and coefficients: csrc/host_finalize.h) ------
static void mass_cheb_kernels(fdal_ctx *c, MassCheb &mc, CgWs &w, const DevCsr &M, const double *invdiag,
                              const double *b, double *x) {
  // iterate ping-pongs between w.x and w.v, d lives in w.p; the last step writes x
  const int64_t n = w.n;
  double *cur = w.x;
  k_cheb_zero<<<grid_elems(c, std::max<int64_t>(n, 1)), kBlock, 0, c->stream>>>(n, b, invdiag, mc.coef[1], w.p,
                                                                                 mc.its == 1 ? x : cur);
  c->launches++;
  for (int k = 1; k < mc.its; ++k) {
    double *dst = k + 1 == mc.its ? x : (cur == w.x ? w.v : w.x);
    EpiCheb<false> e{b, invdiag, cur, w.p, dst, mc.coef[2 * (size_t)k], mc.coef[2 * (size_t)k + 1], 0};
    spmv(c, M, cur, e);
    cur = dst;
  }
}
// x = M^-1 b, Chebyshev mode 1.  The graph is pointer-bound, so the operands are staged through w.bin / w.z
// (two m-vector copies) like the fixed-count PCG does.
static void mass_cheb_solve1(fdal_ctx *c, MassCheb &mc, CgWs &w, const DevCsr &M, const double *invdiag,
                             const double *b, double *x) {
  if (graphs_enabled(c) && mc.graph_ok && !stream_is_capturing(c)) {
    if (!mc.exec)
      mc.graph_ok = capture_graph(c, [&]() { mass_cheb_kernels(c, mc, w, M, invdiag, w.bin, w.z); }, &mc.exec,
                                  &mc.exec_nodes);
    if (mc.exec) {
      dcopy(c, w.n, b, w.bin);
      if (cudaGraphLaunch(mc.exec, c->stream) != cudaSuccess && !c->fail) c->fail = FDAL_ERR_CUDA;
      c->launches += mc.exec_nodes;
      c->graph_launches++;
      dcopy(c, w.n, w.z, x);
      return;
    }
  }
  mass_cheb_kernels(c, mc, w, M, invdiag, b, x);
}
// y = a * M^-repeat b (+ add), Chebyshev mode 2: one persistent kernel
static void mass_cheb_solve2(fdal_ctx *c, MassCheb &mc, const DevCsr &M, const double *invdiag, int repeat, double a,
                             const double *b, const double *add, double *y) {
  cudaMemsetAsync(mc.bar, 0, sizeof(unsigned int), c->stream);  // arrivals only: the give-up flag bar[1] is sticky
  k_mass_cheb_grid<<<mc.grid, kBlock, mc.smem, c->stream>>>(M.d, invdiag, mc.d_coef, mc.its, repeat, a, b, add, y,
                                                            mc.xbuf, mc.bar, mc.rpc, mc.nnz_cap);
  c->launches++;
  launch_check(c);
}
// Decide whether (and how) the exact solve with M runs as a Chebyshev iteration: bounds from the
// calibration CG, iteration count from the Chebyshev error bound, then a check of the true residual
// on the calibration right-hand side (still in w.bin).  Anything unexpected keeps the Jacobi-PCG.
static int mass_cheb_setup(fdal_ctx *c, MassCheb &mc, CgWs &w, const DevCsr &M, const double *invdiag,
                           const CgHistory &h) {
  // FDAL_MASS_CHEB = 0: Jacobi-PCG, 1 (default): one fused kernel per iteration, 2: the persistent kernel.
  // Measured on B200 (profiles/r2_mass_chebyshev.md, elliptic_interface cycle 6, m = 20 609, 58 iterations): one
  // augmented apply 928 us with the Jacobi-PCG (244 launches), 220 us with one kernel per iteration, 267 us with the
  // persistent kernel — a grid barrier through L2 (store, fence, arrive, poll, gather: four dependent L2 round
  // trips, ~4 us) costs more than a CUDA-graph kernel node (~2.9 us), so the graph is the default.
  const char *env = getenv("FDAL_MASS_CHEB");
  const int want = env ? atoi(env) : 1;
  mc.mode = 0;
  if (want <= 0 || w.dist || w.n < 2 || M.use_bsr) return FDAL_OK;
  const int cap = c->cfg.exact_mass_max_its > 0 ? c->cfg.exact_mass_max_its : 300;
  if (!chebyshev_plan(h, cap, &mc.lo, &mc.hi, &mc.its, mc.coef)) return FDAL_OK;
  const int its = mc.its;
  int st;
  int mode = 1;
  if (want >= 2 && M.d.nnz < (1ll << 31)) {
    // persistent kernel: at most one CTA per SM, >= 128 rows each (FDAL_MASS_CHEB_RPC: test knob, so that
    // small problems exercise the grid barrier too); every CTA's slice must fit shared memory
    const int n = (int)w.n;
    const char *rpc_env = getenv("FDAL_MASS_CHEB_RPC");
    const int min_rpc = rpc_env && atoi(rpc_env) > 0 ? atoi(rpc_env) : 128;
    int grid = std::max(1, std::min(c->sms, (n + min_rpc - 1) / min_rpc));
    const int rpc = (n + grid - 1) / grid;
    grid = (n + rpc - 1) / rpc;
    std::vector<int> hrp((size_t)n + 1);
    CU(cudaMemcpyAsync(hrp.data(), M.rp, ((size_t)n + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int nnz_cap = 0;
    for (int j = 0; j < grid; ++j)
      nnz_cap = std::max(nnz_cap, hrp[(size_t)std::min(n, (j + 1) * rpc)] - hrp[(size_t)std::min(n, j * rpc)]);
    nnz_cap = (nnz_cap + 1) & ~1;
    const size_t smem = (size_t)rpc * 4 * sizeof(double) + (size_t)nnz_cap * (sizeof(double) + sizeof(int)) +
                        ((size_t)rpc + 2) * sizeof(int);
    int dev = 0, max_optin = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem <= (size_t)max_optin) {
      // the opt-in maximum, not `smem`: M and Mp may both use the kernel with different slice sizes
      CU(cudaFuncSetAttribute(k_mass_cheb_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, max_optin));
      int per_sm = 0;
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mass_cheb_grid, kBlock, smem));
      if (per_sm >= 1) {  // co-residency of the whole grid: the barrier's precondition
        mc.grid = grid;
        mc.rpc = rpc;
        mc.nnz_cap = nnz_cap;
        mc.smem = smem;
        if ((st = dvec(c, &mc.d_coef, 2 * (int64_t)its))) return st;
        if ((st = dvec(c, &mc.xbuf, 2 * (int64_t)n))) return st;
        if ((st = dmalloc(c, &mc.bar, 2))) return st;
        CU(cudaMemsetAsync(mc.bar, 0, 2 * sizeof(unsigned int), c->stream));
        CU(cudaMemcpyAsync(mc.d_coef, mc.coef.data(), mc.coef.size() * sizeof(double), cudaMemcpyHostToDevice,
                           c->stream));
        CU(cudaStreamSynchronize(c->stream));
        mode = 2;
      }
    }
  }
  // verification on the calibration right-hand side
  mc.mode = mode;
  if (mode == 2)
    mass_cheb_solve2(c, mc, M, invdiag, 1, 1.0, w.bin, nullptr, w.z);
  else
    mass_cheb_kernels(c, mc, w, M, invdiag, w.bin, w.z);
  spmv(c, M, w.z, EpiResid{w.r, w.bin});
  dot(c, w.n, w.r, w.r, w.scal + S_RR, false);
  dot(c, w.n, w.bin, w.bin, w.scal + S_RHO, false);
  double sc[4];
  if ((st = read_scalars(c, w.scal, 4, sc))) return st;
  mc.verified_resid = std::sqrt(std::fabs(sc[S_RR]) / std::max(sc[S_RHO], 1e-300));
  if (mode == 2) {
    unsigned int gave_up = 0;
    CU(cudaMemcpy(&gave_up, mc.bar + 1, sizeof(gave_up), cudaMemcpyDeviceToHost));
    if (gave_up) mc.verified_resid = std::numeric_limits<double>::infinity();
  }
  if (!(mc.verified_resid <= 2e-14)) {
    set_err(c, "warning: the Chebyshev mass solve (%d iterations on [%.4g, %.4g]) left a relative residual of %.2e: "
               "keeping the Jacobi-PCG", its, mc.lo, mc.hi, mc.verified_resid);
    mc.mode = 0;
  }
  return FDAL_OK;
}
// x = M^-1 b through whichever form was selected at fdal_finalize
static void mass_solve(fdal_ctx *c, MassCheb &mc, CgWs &w, const DevCsr &M, const double *invdiag, int pcg_its,
                       const double *b, double *x) {
  if (mc.mode == 2)
    mass_cheb_solve2(c, mc, M, invdiag, 1, 1.0, b, nullptr, x);
  else if (mc.mode == 1)
    mass_cheb_solve1(c, mc, w, M, invdiag, b, x);
  else
    mass_solve_fixed(c, w, M, invdiag, pcg_its, b, x);
}

// ------------------------------------------------------------------ operators of the path
static bool is_stokes(const fdal_ctx *c) {
  return c->cfg.kind == FDAL_KIND_STOKES || c->cfg.kind == FDAL_KIND_STOKES_DIAG_MINRES;
}
static bool is_elliptic(const fdal_ctx *c) {
  return c->cfg.kind == FDAL_KIND_ELLIPTIC_IDEAL || c->cfg.kind == FDAL_KIND_ELLIPTIC_MODIFIED;
}

// y = a * invW x (+ add)   (K4 / K5).  x and y must be distinct buffers (and not t_mw).
static void apply_winv_scaled(fdal_ctx *c, double a, const double *x, double *y, const double *add = nullptr) {
  const int64_t m = c->m;
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    diag_scale(c, m, a, c->d_winv, x, y);
    if (add) axpby(c, m, 1.0, add, 1.0, y);
    return;
  }
  const int repeat = c->cfg.winv_mode == FDAL_WINV_EXACT_M ? 1 : 2;
  if (c->d_winv_dense) {
    k_gemv_axpb<<<(int)((m * 32 + kBlock - 1) / kBlock), kBlock, 0, c->stream>>>((int)m, c->d_winv_dense, x, a, add, y);
    c->launches++;
    return;
  }
  if (c->mass_cta_ws) {
    // whole fixed-count PCG (both applications of M^-1 for W = M^2) in one CTA
    const int threads = (int)std::min<int64_t>(kMassCtaThreads, ((m + 31) / 32) * 32);
    if (m <= kMassCtaSmemRows) {
      const size_t sm = (size_t)4 * m * sizeof(double);  // opt-in above 48 KB: set in fdal_finalize, per device
      k_mass_pcg_cta<true><<<1, threads, sm, c->stream>>>(c->dmat[FDAL_MAT_M].d, c->d_m_invdiag, c->mass_its_m, repeat,
                                                          a, x, add, y, c->mass_cta_ws);
    } else {
      k_mass_pcg_cta<false><<<1, threads, 0, c->stream>>>(c->dmat[FDAL_MAT_M].d, c->d_m_invdiag, c->mass_its_m,
                                                          repeat, a, x, add, y, c->mass_cta_ws);
    }
    c->launches++;
    return;
  }
  if (c->mcheb_m.mode == 2) {  // whole solve(s), scaling and the add in one persistent kernel
    mass_cheb_solve2(c, c->mcheb_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, repeat, a, x, add, y);
    return;
  }
  if (repeat == 1) {
    mass_solve(c, c->mcheb_m, c->cgmass_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, c->mass_its_m, x, y);
  } else {
    mass_solve(c, c->mcheb_m, c->cgmass_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, c->mass_its_m, x, c->t_mw);
    mass_solve(c, c->mcheb_m, c->cgmass_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, c->mass_its_m, c->t_mw, y);
  }
  dscale(c, m, a, y);
  if (add) axpby(c, m, 1.0, add, 1.0, y);
}
// Mp_inv (stokes_immersed_boundary.cc:931-963)
static int apply_mp_inv(fdal_ctx *c, const double *x, double *y) {
  if (c->cfg.mp_inv_mode == FDAL_MPINV_EXACT) {
    mass_solve(c, c->mcheb_p, c->cgmass_p, c->dmat[FDAL_MAT_MP], c->d_mp_invdiag, c->mass_its_p, x, y);
    return FDAL_OK;
  }
  OpFn op = [&](const double *in, double *out, double *d) { spmv(c, c->dmat[FDAL_MAT_MP], in, EpiDotX{out, in}, d); };
  OpFn pr = [&](const double *r, double *z, double *d) {
    mass_prec(c, c->d_mp_lumped, c->n1, r, z, d, c->cgmass_p.dist);
  };
  int its = 0;
  int st = cg_solve(c, c->cgmass_p, op, pr, c->cfg.mass, x, y, &its, FDAL_ERR_MASS_NO_CONVERGENCE, true);
  c->its_mass += its;
  if (st && !c->fail) c->fail = st;
  return st;
}

// phase 1 of the fused augmented apply: t = a * invW (C x [- M x1m]) + add ; y1 = C x
static void couple_phase1(fdal_ctx *c, const double *x, double a, const double *add, double *y1, double *t) {
  const DevCsr &C = c->dmat[FDAL_MAT_C];
  if (c->nranks > 1) {
    // C holds this rank's columns: partial sums, all-reduced over the m multiplier rows
    double *w = (y1 && c->cfg.winv_mode != FDAL_WINV_DIAG) ? y1 : c->t_m1;
    spmv(c, C, x, EpiAssign{w, 1.0});
    allreduce_vec(c, w, (size_t)c->m);
    if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
      k_couple_elem<<<grid_elems(c, c->m), kBlock, 0, c->stream>>>((int)c->m, w, c->d_winv, a, add, y1, t);
      c->launches++;
    } else {
      if (y1 && w != y1) dcopy(c, c->m, w, y1);
      apply_winv_scaled(c, a, w, t, add);
    }
    return;
  }
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    if (C.use_merge) {  // few, long rows: merge-path C x, then the element-wise coupling epilogue
      spmv(c, C, x, EpiAssign{c->t_m1, 1.0});
      k_couple_elem<<<grid_elems(c, c->m), kBlock, 0, c->stream>>>((int)c->m, c->t_m1, c->d_winv, a, add, y1, t);
      c->launches++;
    } else {
      spmv(c, C, x, EpiCouple{t, c->d_winv, a, add, y1});
    }
  } else {
    spmv(c, C, x, EpiCouple{c->t_m1, nullptr, 1.0, nullptr, y1});
    apply_winv_scaled(c, a, c->t_m1, t, add);
  }
}

// Aug.vmult (K3): y = A x + gamma Ct invW C x  [+ gamma_gd Bt Mp^-1 B x]; optional dot x.y
static void apply_aug11(fdal_ctx *c, const double *x, double *y, double *dot_out) {
  const DevCsr &A = c->dmat[FDAL_MAT_A];
  const bool gd = is_stokes(c) && c->cfg.grad_div_in_operator;
  if (c->cfg.aug_explicit) {
    if (dot_out && !gd)
      spmv(c, A, x, EpiDotX{y, x}, dot_out);
    else
      spmv(c, A, x, EpiAssign{y, 1.0});
  } else if (c->overlap_mass && c->mass_cta_ws && !c->d_winv_dense && c->nranks == 1 && !gd) {
    // exact W^-1: the mass solve is ONE CTA for ~90 us.  Fork: y = A x on the second stream
    // fills the other 147 SMs meanwhile; join: y += Ct t (+ fused x.y) over the few rows of Ct.
    // C x first, then fork: the one-CTA mass kernel is enqueued BEFORE the grid-filling A x
    // (whose persistent CTAs would otherwise hold every SM until they finish)
    spmv(c, c->dmat[FDAL_MAT_C], x, EpiCouple{c->t_m1, nullptr, 1.0, nullptr, nullptr});
    cudaEventRecord(c->ev_fork, c->stream);
    apply_winv_scaled(c, c->cfg.gamma, c->t_m1, c->t_m0);
    cudaStreamWaitEvent(c->stream2, c->ev_fork, 0);
    cudaStream_t main_stream = c->stream;
    c->stream = c->stream2;
    spmv(c, A, x, EpiAssign{y, 1.0});
    c->stream = main_stream;
    cudaEventRecord(c->ev_join, c->stream2);
    cudaStreamWaitEvent(c->stream, c->ev_join, 0);
    if (dot_out)
      spmv(c, c->dmat[FDAL_MAT_CT], c->t_m0, EpiAddDotX{y, x}, dot_out);
    else
      spmv(c, c->dmat[FDAL_MAT_CT], c->t_m0, EpiAdd{y, 1.0});
  } else {
    couple_phase1(c, x, c->cfg.gamma, nullptr, nullptr, c->t_m0);
    if (dot_out && !gd)
      spmv2(c, A, x, c->dmat[FDAL_MAT_CT], c->t_m0, EpiDotX{y, x}, dot_out);
    else
      spmv2(c, A, x, c->dmat[FDAL_MAT_CT], c->t_m0, EpiAssign{y, 1.0});
  }
  if (gd) {
    spmv(c, c->dmat[FDAL_MAT_B], x, EpiAssign{c->t_p0, 1.0});
    apply_mp_inv(c, c->t_p0, c->t_p1);
    spmv(c, c->dmat[FDAL_MAT_BT], c->t_p1, EpiAdd{y, c->cfg.gamma_grad_div});
    if (dot_out) dot(c, c->n0, x, y, dot_out, c->nranks > 1);
  }
}
// A22_aug = A2 + gamma_2 M invW M (elliptic_interface.cc:810)
static void apply_aug22(fdal_ctx *c, const double *x, double *y, double *dot_out) {
  const DevCsr &M = c->dmat[FDAL_MAT_M];
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    spmv(c, M, x, EpiCouple{c->t_m0, c->d_winv, c->cfg.gamma2, nullptr, nullptr});
  } else {
    spmv(c, M, x, EpiAssign{c->t_m1, 1.0});
    apply_winv_scaled(c, c->cfg.gamma2, c->t_m1, c->t_m0);
  }
  if (dot_out)
    spmv2(c, c->dmat[FDAL_MAT_A2], x, M, c->t_m0, EpiDotX{y, x}, dot_out);
  else
    spmv2(c, c->dmat[FDAL_MAT_A2], x, M, c->t_m0, EpiAssign{y, 1.0});
}
static void apply_aug(fdal_ctx *c, int which, const double *x, double *y, double *dot_out) {
  if (which == FDAL_AMG_A11)
    apply_aug11(c, x, y, dot_out);
  else
    apply_aug22(c, x, y, dot_out);
}

// elliptic 2x2 augmented block [[A11g,A12g],[A21g,A22g]] applied to [x0;x1]:
//   w = C x0 - M x1 ; tw = invW w ; y0 = A1 x0 + g1 Ct tw ; y1 = A2 x1 - g2 M tw
// (the four LinearOperator blocks of elliptic_interface.cc:807-813 share tw)
static void elliptic_w(fdal_ctx *c, const double *x0, const double *x1, double *w) {
  spmv(c, c->dmat[FDAL_MAT_C], x0, EpiAssign{w, 1.0});
  allreduce_vec(c, w, (size_t)c->m);
  spmv(c, c->dmat[FDAL_MAT_M], x1, EpiAdd{w, -1.0});
}
static void apply_aug_block(fdal_ctx *c, const double *x, double *y, double *dot_out) {
  const double *x0 = x, *x1 = x + c->n0;
  double *y0 = y, *y1 = y + c->n0;
  elliptic_w(c, x0, x1, c->t_m1);
  apply_winv_scaled(c, 1.0, c->t_m1, c->t_m2);  // tw (exact modes use t_m2 internally: see below)
  axpby(c, c->m, c->cfg.gamma, c->t_m2, 0.0, c->t_m0);
  spmv2(c, c->dmat[FDAL_MAT_A], x0, c->dmat[FDAL_MAT_CT], c->t_m0, EpiAssign{y0, 1.0});
  axpby(c, c->m, -c->cfg.gamma2, c->t_m2, 0.0, c->t_m0);
  spmv2(c, c->dmat[FDAL_MAT_A2], x1, c->dmat[FDAL_MAT_M], c->t_m0, EpiAssign{y1, 1.0});
  if (dot_out) dot(c, c->cgblk.dist ? c->cgblk.n_dot : c->n0 + c->n1, x, y, dot_out, c->cgblk.dist);
}

// AA.vmult / system_operator.vmult (a7)
static void apply_system(fdal_ctx *c, const double *x, double *y) {
  const double *x0 = x, *x1 = x + c->n0, *x2 = x + c->n0 + c->n1;
  double *y0 = y, *y1 = y + c->n0, *y2 = y + c->n0 + c->n1;
  const DevCsr &A = c->dmat[FDAL_MAT_A], &Ct = c->dmat[FDAL_MAT_CT];
  switch (c->cfg.kind) {
    case FDAL_KIND_LAPLACE:
      // y0 = A x0 + Ct (gamma invW C x0 + x1) ; y1 = C x0
      if (c->cfg.aug_explicit) {
        spmv(c, c->dmat[FDAL_MAT_C], x0, EpiAssign{y1, 1.0});
        allreduce_vec(c, y1, (size_t)c->m);
        spmv2(c, A, x0, Ct, x1, EpiAssign{y0, 1.0});
      } else {
        couple_phase1(c, x0, c->cfg.gamma, x1, y1, c->t_m0);
        spmv2(c, A, x0, Ct, c->t_m0, EpiAssign{y0, 1.0});
      }
      break;
    case FDAL_KIND_STOKES:
    case FDAL_KIND_STOKES_DIAG_MINRES:
      if (c->cfg.aug_explicit) {
        spmv(c, c->dmat[FDAL_MAT_C], x0, EpiAssign{y2, 1.0});
        allreduce_vec(c, y2, (size_t)c->m);
        spmv2(c, A, x0, Ct, x2, EpiAssign{y0, 1.0});
      } else {
        couple_phase1(c, x0, c->cfg.gamma, x2, y2, c->t_m0);
        spmv2(c, A, x0, Ct, c->t_m0, EpiAssign{y0, 1.0});
      }
      spmv(c, c->dmat[FDAL_MAT_BT], x1, EpiAdd{y0, 1.0});
      spmv(c, c->dmat[FDAL_MAT_B], x0, EpiAssign{y1, 1.0});
      if (c->cfg.grad_div_in_operator) {
        apply_mp_inv(c, y1, c->t_p1);
        spmv(c, c->dmat[FDAL_MAT_BT], c->t_p1, EpiAdd{y0, c->cfg.gamma_grad_div});
      }
      break;
    default: {
      // w = C x0 - M x1 ; tw = invW w
      // y0 = A1 x0 + Ct (g1 tw + x2) ; y1 = A2 x1 - M (g2 tw + x2) ; y2 = w
      elliptic_w(c, x0, x1, y2);
      apply_winv_scaled(c, 1.0, y2, c->t_m2);
      axpby(c, c->m, 1.0, x2, 0.0, c->t_m0);
      axpby(c, c->m, c->cfg.gamma, c->t_m2, 1.0, c->t_m0);
      spmv2(c, A, x0, Ct, c->t_m0, EpiAssign{y0, 1.0});
      axpby(c, c->m, -1.0, x2, 0.0, c->t_m0);
      axpby(c, c->m, -c->cfg.gamma2, c->t_m2, 1.0, c->t_m0);
      spmv2(c, c->dmat[FDAL_MAT_A2], x1, c->dmat[FDAL_MAT_M], c->t_m0, EpiAssign{y1, 1.0});
    } break;
  }
}

// Aug_inv = inverse_operator(Aug, SolverCG, AMG) (a8)
static int apply_aug_inv(fdal_ctx *c, int which, const double *b, double *x, int *its) {
  CgWs &w = which == FDAL_AMG_A11 ? c->cg11 : c->cg22;
  OpFn op = [c, which](const double *in, double *out, double *d) { apply_aug(c, which, in, out, d); };
  OpFn pr;
  if (c->cfg.inner_prec == FDAL_PREC_AMG)
    pr = [c, which](const double *r, double *z, double *d) { vcycle(c, c->amg[which], r, z, d); };
  else
    pr = [c, &w](const double *r, double *z, double *d) {
      dcopy(c, w.n, r, z);
      dot(c, w.dist ? w.n_dot : w.n, r, z, d, w.dist);
    };
  // the only non-capturable body: the no-grad-div Stokes operator nests a host-checked Mp CG
  const bool capturable = !(which == FDAL_AMG_A11 && is_stokes(c) && c->cfg.grad_div_in_operator &&
                            c->cfg.mp_inv_mode == FDAL_MPINV_CG_LUMPED);
  int st = cg_solve(c, w, op, pr, c->cfg.inner, b, x, its, FDAL_ERR_INNER_NO_CONVERGENCE, capturable);
  if (which == FDAL_AMG_A11)
    c->its_a11 += *its;
  else
    c->its_a22 += *its;
  c->n_inner_solves++;
  if (st && !c->fail) c->fail = st;
  return st;
}

// the five P.vmult (a1-a5)
static int apply_prec(fdal_ctx *c, const double *u, double *v) {
  const double *u0 = u, *u1 = u + c->n0, *u2 = u + c->n0 + c->n1;
  double *v0 = v, *v1 = v + c->n0, *v2 = v + c->n0 + c->n1;
  const double g = c->cfg.gamma;
  const DevCsr &Ct = c->dmat[FDAL_MAT_CT];
  int its = 0;
  switch (c->cfg.kind) {
    case FDAL_KIND_LAPLACE:
      apply_winv_scaled(c, -g, u1, v1);
      spmv(c, Ct, v1, EpiResid{c->t_n0, u0});
      apply_aug_inv(c, FDAL_AMG_A11, c->t_n0, v0, &its);
      break;
    case FDAL_KIND_STOKES:
      apply_winv_scaled(c, -g, u2, v2);
      apply_mp_inv(c, u1, v1);
      dscale(c, c->n1, -c->cfg.gamma_grad_div, v1);
      spmv(c, c->dmat[FDAL_MAT_BT], v1, EpiResid{c->t_n0, u0});
      spmv(c, Ct, v2, EpiAdd{c->t_n0, -1.0});
      apply_aug_inv(c, FDAL_AMG_A11, c->t_n0, v0, &its);
      break;
    case FDAL_KIND_STOKES_DIAG_MINRES:
      apply_winv_scaled(c, g, u2, v2);
      apply_mp_inv(c, u1, v1);
      dscale(c, c->n1, c->cfg.gamma_grad_div, v1);
      apply_aug_inv(c, FDAL_AMG_A11, u0, v0, &its);
      break;
    case FDAL_KIND_ELLIPTIC_IDEAL: {
      apply_winv_scaled(c, -g, u2, v2);
      double *uu = c->t_N0;  // [n0 | n1]
      spmv(c, Ct, v2, EpiResid{uu, u0});
      dcopy(c, c->n1, u1, uu + c->n0);
      spmv(c, c->dmat[FDAL_MAT_M], v2, EpiAdd{uu + c->n0, 1.0});
      OpFn op = [c](const double *in, double *out, double *d) { apply_aug_block(c, in, out, d); };
      OpFn pr;
      if (c->cfg.inner_prec == FDAL_PREC_AMG)
        pr = [c](const double *r, double *z, double *d) {
          vcycle(c, c->amg[0], r, z, nullptr);
          vcycle(c, c->amg[1], r + c->n0, z + c->n0, nullptr);
          dot(c, c->cgblk.dist ? c->cgblk.n_dot : c->n0 + c->n1, r, z, d, c->cgblk.dist);
        };
      else
        pr = [c](const double *r, double *z, double *d) {
          dcopy(c, c->n0 + c->n1, r, z);
          dot(c, c->cgblk.dist ? c->cgblk.n_dot : c->n0 + c->n1, r, z, d, c->cgblk.dist);
        };
      int st = cg_solve(c, c->cgblk, op, pr, c->cfg.inner, uu, v, &its, FDAL_ERR_INNER_NO_CONVERGENCE, true);
      c->its_a11 += its;
      c->n_inner_solves++;
      if (st && !c->fail) c->fail = st;
    } break;
    case FDAL_KIND_ELLIPTIC_MODIFIED: {
      apply_winv_scaled(c, -g, u2, v2);
      // d1 = A22inv (u2 + M d2)
      dcopy(c, c->n1, u1, c->t_p0);
      spmv(c, c->dmat[FDAL_MAT_M], v2, EpiAdd{c->t_p0, 1.0});
      apply_aug_inv(c, FDAL_AMG_A22, c->t_p0, v1, &its);
      // d0 = A11inv (u + gamma Ct invW M d1 - Ct d2) = A11inv (u + Ct (gamma invW M d1 - d2))
      spmv(c, c->dmat[FDAL_MAT_M], v1, EpiAssign{c->t_m1, 1.0});
      // t = d2 - gamma invW M d1  =>  u0 - Ct t is the bracket above
      apply_winv_scaled(c, -g, c->t_m1, c->t_m0);
      axpby(c, c->m, 1.0, v2, 1.0, c->t_m0);
      spmv(c, Ct, c->t_m0, EpiResid{c->t_n0, u0});
      apply_aug_inv(c, FDAL_AMG_A11, c->t_n0, v0, &its);
    } break;
  }
  return c->fail;
}

}  // namespace fdal

// ====================================================================== outer solvers
namespace fdal {

static void record(fdal_solve_info *info, double res) {
  if (info->n_history < FDAL_MAX_HISTORY) info->residual_history[info->n_history++] = res;
}

// SolverFGMRES (a12): batched classical Gram-Schmidt with one
// re-orthogonalisation pass; Hessenberg / Givens on the host (O(restart^2)).
static int fgmres(fdal_ctx *c, const double *b, double *x, fdal_solve_info *info) {
  const int64_t N = c->N;
  const int64_t Nd = c->n_dot_outer;  // entries this rank contributes to dots
  const bool dist = c->nranks > 1;
  const int mb = c->cfg.restart;
  std::vector<double> H((size_t)(mb + 1) * mb, 0.0), g(mb + 1), cs(mb), sn(mb), y(mb), hh(2 * (mb + 2));
  ControlState ctl;
  ctl.c = c->cfg.outer;
  int acc = 0, state = ST_ITERATE, st;
  auto Vj = [&](int j) { return c->V + (size_t)j * N; };
  auto Zj = [&](int j) { return c->Z + (size_t)j * N; };
  do {
    apply_system(c, x, Vj(0));
    axpby(c, N, 1.0, b, -1.0, Vj(0));
    dot(c, Nd, Vj(0), Vj(0), c->d_scal, dist);
    double rr;
    if ((st = read_scalars(c, c->d_scal, 1, &rr))) return st;
    double res = std::sqrt(rr);
    if (acc == 0) {
      info->initial_residual = res;
      record(info, res);
    }
    state = control_check(ctl, acc, res);
    if (state != ST_ITERATE) break;
    k_scale_by_inv<<<grid_elems(c, N), kBlock, 0, c->stream>>>(N, Vj(0), c->d_scal, 1, Vj(0));
    c->launches++;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = res;
    int j = 0;
    for (; j < mb && state == ST_ITERATE; ++j) {
      apply_prec(c, Vj(j), Zj(j));
      if (c->fail) break;
      double *w = Vj(j + 1);
      apply_system(c, Zj(j), w);
      const int nv = j + 1;
      double *h1 = c->d_h, *h2 = c->d_h + (mb + 2);
      const int gd = grid_elems(c, N);
      k_multidot<<<gd, kBlock, 0, c->stream>>>(N, Nd, w, c->V, N, nv, 0, reducer(c, h1, dist));
      finish_scalar(c, h1, (size_t)nv, dist);
      k_multiaxpy<<<gd, kBlock, 0, c->stream>>>(N, w, c->V, N, nv, h1, -1.0);
      k_multidot<<<gd, kBlock, 0, c->stream>>>(N, Nd, w, c->V, N, nv, 0, reducer(c, h2, dist));
      finish_scalar(c, h2, (size_t)nv, dist);
      k_multiaxpy<<<gd, kBlock, 0, c->stream>>>(N, w, c->V, N, nv, h2, -1.0);
      dot(c, Nd, w, w, h2 + nv, dist);
      k_scale_by_inv<<<gd, kBlock, 0, c->stream>>>(N, w, h2 + nv, 1, w);
      c->launches += 5;
      // one read-back per outer iteration: h1[0..nv), h2[0..nv], |w|^2
      CU(cudaMemcpyAsync(c->h_scal, c->d_h, (size_t)(2 * (mb + 2)) * sizeof(double), cudaMemcpyDeviceToHost,
                         c->stream));
      CU(cudaStreamSynchronize(c->stream));
      double *h = &H[(size_t)j * (mb + 1)];
      for (int i = 0; i < nv; ++i) h[i] = c->h_scal[i] + c->h_scal[(mb + 2) + i];
      h[j + 1] = std::sqrt(c->h_scal[(mb + 2) + nv]);
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * h[i] + sn[i] * h[i + 1];
        h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
        h[i] = t;
      }
      const double den = std::hypot(h[j], h[j + 1]);
      cs[j] = h[j] / den;
      sn[j] = h[j + 1] / den;
      h[j] = den;
      h[j + 1] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      res = std::fabs(g[j + 1]);
      ++acc;
      record(info, res);
      state = control_check(ctl, acc, res);
    }
    if (c->fail) state = ST_FAILURE;
    const int k = j;
    for (int i = k - 1; i >= 0; --i) {
      double s = g[i];
      for (int l = i + 1; l < k; ++l) s -= H[(size_t)l * (mb + 1) + i] * y[l];
      y[i] = s / H[(size_t)i * (mb + 1) + i];
    }
    if (k > 0) {
      CU(cudaMemcpyAsync(c->d_y, y.data(), (size_t)k * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      k_multiaxpy<<<grid_elems(c, N), kBlock, 0, c->stream>>>(N, x, c->Z, N, k, c->d_y, 1.0);
      c->launches++;
      CU(cudaStreamSynchronize(c->stream));  // y is reused by the next cycle
    }
    if (c->fail) break;
  } while (state == ST_ITERATE);
  info->outer_iterations = ctl.last_step;
  info->final_residual = ctl.last_value;
  if (c->fail) return c->fail;
  return state == ST_SUCCESS ? FDAL_OK : FDAL_ERR_OUTER_NO_CONVERGENCE;
}

// SolverMinRes (a13)
static int minres(fdal_ctx *c, const double *b, double *x, fdal_solve_info *info) {
  const int64_t N = c->N;
  double *u[3] = {c->mr_u[0], c->mr_u[1], c->mr_u[2]}, *m[3] = {c->mr_m[0], c->mr_m[1], c->mr_m[2]}, *v = c->mr_v;
  double delta[3] = {0, 0, 0}, f[2] = {0, 0}, e[2] = {0, 0};
  double r_l2, r0, tau = 0, cc = 0, s = 0, d_ = 0;
  ControlState ctl;
  ctl.c = c->cfg.outer;
  int j = 1, st;
  auto hdot = [&](const double *a, const double *bb, double *out) -> int {
    dot(c, c->n_dot_outer, a, bb, c->d_scal, c->nranks > 1);
    return read_scalars(c, c->d_scal, 1, out);
  };
  apply_system(c, x, m[0]);
  dcopy(c, N, b, u[1]);
  axpby(c, N, -1.0, m[0], 1.0, u[1]);
  apply_prec(c, u[1], v);
  if ((st = hdot(v, u[1], &delta[1]))) return st;
  r0 = std::sqrt(std::fabs(delta[1]));
  r_l2 = r0;
  dzero(c, N, u[0]);
  dzero(c, N, u[2]);
  for (int i = 0; i < 3; ++i) dzero(c, N, m[i]);
  info->initial_residual = r_l2;
  record(info, r_l2);
  int state = control_check(ctl, 0, r_l2);
  while (state == ST_ITERATE && !c->fail) {
    if (delta[1] != 0)
      dscale(c, N, 1.0 / std::sqrt(delta[1]), v);
    else
      dzero(c, N, v);
    apply_system(c, v, u[2]);
    if (j > 1) axpby(c, N, -std::sqrt(delta[1] / delta[0]), u[0], 1.0, u[2]);
    double gamma;
    if ((st = hdot(u[2], v, &gamma))) return st;
    axpby(c, N, -gamma / std::sqrt(delta[1]), u[1], 1.0, u[2]);
    dcopy(c, N, v, m[0]);
    apply_prec(c, u[2], v);
    if ((st = hdot(v, u[2], &delta[2]))) return st;
    const double sd2 = std::sqrt(std::fabs(delta[2]));
    if (j == 1) {
      d_ = gamma;
      e[1] = sd2;
    }
    if (j > 1) {
      d_ = s * e[0] - cc * gamma;
      e[0] = cc * e[0] + s * gamma;
      f[1] = s * sd2;
      e[1] = -cc * sd2;
    }
    const double d = std::sqrt(d_ * d_ + std::fabs(delta[2]));
    if (j > 1) tau *= s / cc;
    cc = d_ / d;
    tau *= cc;
    s = sd2 / d;
    if (j == 1) tau = r0 * cc;
    axpby(c, N, -e[0], m[1], 1.0, m[0]);
    if (j > 1) axpby(c, N, -f[0], m[2], 1.0, m[0]);
    dscale(c, N, 1.0 / d, m[0]);
    axpby(c, N, tau, m[0], 1.0, x);
    r_l2 *= std::fabs(s);
    record(info, r_l2);
    state = control_check(ctl, j, r_l2);
    ++j;
    std::swap(m[2], m[1]);
    std::swap(m[1], m[0]);
    std::swap(u[0], u[1]);
    std::swap(u[1], u[2]);
    delta[0] = delta[1];
    delta[1] = delta[2];
    f[0] = f[1];
    e[0] = e[1];
  }
  info->outer_iterations = ctl.last_step;
  info->final_residual = ctl.last_value;
  if (c->fail) return c->fail;
  return state == ST_SUCCESS ? FDAL_OK : FDAL_ERR_OUTER_NO_CONVERGENCE;
}

// ------------------------------------------------------------------ finalize helpers
// dense W^-1 of a small multiplier space (opt-in): Gauss-Jordan inverse of M on the device (same
// kernels as the coarsest AMG operator), squared for W = M^2.  2 MB for the 514 multipliers of
// configs[1]: it stays in L2 and one GEMV replaces the two ~45 us single-CTA mass solves.
static const int64_t kDenseWinvMaxRows = 4096;
static int build_dense_winv(fdal_ctx *c) {
  const DevCsr &M = c->dmat[FDAL_MAT_M];
  const int n = (int)c->m;
  double *aug = nullptr, *col = nullptr, *pval = nullptr, *minv = nullptr;
  int *prow = nullptr, *sing = nullptr;
  int st;
  if ((st = dvec(c, &aug, (int64_t)n * 2 * n))) return st;
  if ((st = dvec(c, &col, n))) return st;
  if ((st = dvec(c, &pval, 1))) return st;
  if ((st = dmalloc(c, &prow, 1))) return st;
  if ((st = dmalloc(c, &sing, 1))) return st;
  if ((st = dmalloc(c, &minv, (size_t)n * n))) return st;
  CU(cudaMemsetAsync(sing, 0, sizeof(int), c->stream));
  const int tb = 256;
  k_dense_from_csr<<<(n + tb - 1) / tb, tb, 0, c->stream>>>(n, M.rp, M.ci, M.v, aug);
  const int gc = (2 * n + tb - 1) / tb;
  for (int k = 0; k < n; ++k) {
    k_gj_pivot<<<1, kBlock, 0, c->stream>>>(n, k, aug, prow, pval, sing);
    k_gj_swap_scale<<<gc, tb, 0, c->stream>>>(n, k, aug, prow, pval);
    k_gj_save_col<<<(n + tb - 1) / tb, tb, 0, c->stream>>>(n, k, aug, col);
    k_gj_eliminate<<<dim3(gc, n), tb, 0, c->stream>>>(n, k, aug, col);
  }
  k_gj_extract<<<dim3((n + tb - 1) / tb, n), tb, 0, c->stream>>>(n, aug, minv);
  int hs = 0;
  CU(cudaMemcpyAsync(&hs, sing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  cudaFree(aug);
  c->allocs.erase(std::find(c->allocs.begin(), c->allocs.end(), (void *)aug));
  if (hs) {
    set_err(c, "immersed mass matrix is singular");
    return FDAL_ERR_INVALID;
  }
  if (c->cfg.winv_mode == FDAL_WINV_EXACT_M) {
    c->d_winv_dense = minv;
    return FDAL_OK;
  }
  if ((st = dmalloc(c, &c->d_winv_dense, (size_t)n * n))) return st;
  k_dense_square<<<dim3((n + tb - 1) / tb, n), tb, 0, c->stream>>>(n, minv, c->d_winv_dense);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  cudaFree(minv);
  c->allocs.erase(std::find(c->allocs.begin(), c->allocs.end(), (void *)minv));
  return FDAL_OK;
}
static int invert_coarse(fdal_ctx *c, Amg &g) {
  AmgLevel &C = g.lev.back();
  const int n = g.cn_global;  // the coarsest operator is replicated on every rank
  double *aug = nullptr, *col = nullptr, *pval = nullptr;
  int *prow = nullptr, *sing = nullptr;
  int st;
  if ((st = dvec(c, &aug, (int64_t)n * 2 * n))) return st;
  if ((st = dvec(c, &col, n))) return st;
  if ((st = dvec(c, &pval, 1))) return st;
  if ((st = dmalloc(c, &prow, 1))) return st;
  if ((st = dmalloc(c, &sing, 1))) return st;
  if ((st = dmalloc(c, &g.cinv, (size_t)n * n))) return st;
  CU(cudaMemsetAsync(sing, 0, sizeof(int), c->stream));
  const int tb = 256;
  k_dense_from_csr<<<(n + tb - 1) / tb, tb, 0, c->stream>>>(n, C.A.rp, C.A.ci, C.A.v, aug);
  const int gc = (2 * n + tb - 1) / tb;
  for (int k = 0; k < n; ++k) {
    k_gj_pivot<<<1, kBlock, 0, c->stream>>>(n, k, aug, prow, pval, sing);
    k_gj_swap_scale<<<gc, tb, 0, c->stream>>>(n, k, aug, prow, pval);
    k_gj_save_col<<<(n + tb - 1) / tb, tb, 0, c->stream>>>(n, k, aug, col);
    k_gj_eliminate<<<dim3(gc, n), tb, 0, c->stream>>>(n, k, aug, col);
  }
  k_gj_extract<<<dim3((n + tb - 1) / tb, n), tb, 0, c->stream>>>(n, aug, g.cinv);
  int hs = 0;
  CU(cudaMemcpyAsync(&hs, sing, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  if (hs) {
    set_err(c, "coarsest AMG operator is singular");
    return FDAL_ERR_INVALID;
  }
  // the 2n^2 scratch is the largest setup allocation: release it now
  cudaFree(aug);
  c->allocs.erase(std::find(c->allocs.begin(), c->allocs.end(), (void *)aug));
  return FDAL_OK;
}

static int prepare_amg(fdal_ctx *c, Amg &g, bool dist) {
  int st;
  const int nl = (int)g.lev.size();
  g.dist = dist;
  if (!dist)
    g.rep_from = 0;
  else if (g.rep_from < 0)
    g.rep_from = nl - 1;  // only the coarsest operator is replicated
  if (dist && (g.rep_from < 1 || g.rep_from > nl - 1)) {
    set_err(c, "a partitioned AMG hierarchy needs at least one partitioned and one replicated level (replicated from %d of %d)",
            g.rep_from, nl);
    return FDAL_ERR_INVALID;
  }
  for (int l = 0; l < nl; ++l) {
    AmgLevel &L = g.lev[l];
    if (!L.hA.set) {
      set_err(c, "AMG level %d was never set", l);
      return FDAL_ERR_STATE;
    }
    const bool coarsest = (l == nl - 1);
    const bool part = dist && l < g.rep_from;  // this level's rows are a partition
    if (L.hA.owned_cols() != L.hA.nr) {
      set_err(c, "AMG level %d: A is %lld x %lld (owned columns), must be square", l, (long long)L.hA.nr,
              (long long)L.hA.owned_cols());
      return FDAL_ERR_SHAPE;
    }
    if (!coarsest && (L.degree < 1 || !(L.lmax > 0.0) || !(L.ratio > 1.0))) {
      set_err(c, "AMG level %d: Chebyshev needs degree >= 1, lambda_max > 0, eig_ratio > 1 (got %d, %g, %g)", l, L.degree,
              L.lmax, L.ratio);
      return FDAL_ERR_INVALID;
    }
    if (coarsest) g.cn_global = (int)L.hA.nr;
    L.n = (int)L.hA.nr;
    if (dist && l == g.rep_from) {
      if (g.c_hi < 0) {
        set_err(c, "fdal_amg_set_coarse_range was not called for the first replicated level");
        return FDAL_ERR_STATE;
      }
      if (g.c_lo < 0 || g.c_hi < g.c_lo || g.c_hi > L.hA.nr) {
        set_err(c, "AMG level %d: owned range [%lld, %lld) outside [0, %lld]", l, (long long)g.c_lo, (long long)g.c_hi,
                (long long)L.hA.nr);
        return FDAL_ERR_SHAPE;
      }
    }
    if ((st = upload_csr(c, L.hA, L.A, (l == 0 && !coarsest && &g == &c->amg[0]) ? c->cfg.block_size : 1))) return st;
    L.A.dist_rows = part;
    if (!coarsest) {
      if (!L.hP.set) {
        set_err(c, "AMG level %d has no prolongator", l);
        return FDAL_ERR_STATE;
      }
      const AmgLevel &Nx = g.lev[l + 1];
      if (L.hP.nr != L.hA.nr || L.hP.owned_cols() != Nx.hA.nr) {
        set_err(c, "AMG level %d: P is %lld x %lld (owned columns), expected %lld x %lld", l, (long long)L.hP.nr,
                (long long)L.hP.owned_cols(), (long long)L.hA.nr, (long long)Nx.hA.nr);
        return FDAL_ERR_SHAPE;
      }
      if ((st = upload_csr(c, L.hP, L.P))) return st;
      if (!L.hR.set) {
        if (part) {
          set_err(c, "AMG level %d: a partitioned hierarchy needs an explicit R with its halo plan", l);
          return FDAL_ERR_STATE;
        }
        host_transpose(L.hP, L.hR);
      }
      const int64_t r_rows = (dist && l + 1 == g.rep_from) ? (g.c_hi - g.c_lo) : Nx.hA.nr;
      if (L.hR.nr != r_rows || L.hR.owned_cols() != L.hA.nr) {
        set_err(c, "AMG level %d: R is %lld x %lld (owned columns), expected %lld x %lld", l, (long long)L.hR.nr,
                (long long)L.hR.owned_cols(), (long long)r_rows, (long long)L.hA.nr);
        return FDAL_ERR_SHAPE;
      }
      if ((st = upload_csr(c, L.hR, L.R))) return st;
      L.P.dist_rows = part;
      L.R.dist_rows = part && l + 1 < g.rep_from;
      if ((st = dvec(c, &L.invd, L.n))) return st;
      if (!L.h_invd.empty()) {
        CU(cudaMemcpyAsync(L.invd, L.h_invd.data(), (size_t)L.n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      } else {
        k_inv_diag_from_csr<<<std::max(1, (L.n + 255) / 256), 256, 0, c->stream>>>(L.n, L.A.rp, L.A.ci, L.A.v, L.invd);
      }
      if ((st = dvec(c, &L.r, L.n))) return st;
      if ((st = dvec(c, &L.d, L.n))) return st;
    }
    if ((st = dvec(c, &L.xa, L.n))) return st;
    if ((st = dvec(c, &L.xb, L.n))) return st;
    if ((st = dvec(c, &L.b, L.n))) return st;
    CU(cudaStreamSynchronize(c->stream));
    // host copies are no longer needed
    const int64_t keep_nr = L.hA.nr;
    L.hA = HostCsr();
    L.hA.set = true;
    L.hA.nr = keep_nr;
    L.hP = HostCsr();
    L.hR = HostCsr();
  }
  if (dist) {
    if ((st = dvec(c, &g.rown, g.c_hi - g.c_lo))) return st;
    if (c->p2p) {
      // all-gather channel of the first replicated level: rank q's rows land at its offset
      // (the peers' offsets are exchanged in comm_build)
      Chan ch;
      ch.bcast = true;
      ch.cap = g.lev[(size_t)g.rep_from].n;
      ch.src_off.assign((size_t)c->nranks, -1);
      ch.src_off[(size_t)c->rank] = g.c_lo;
      for (int q = 0; q < c->nranks; ++q) ch.nb.push_back(q);
      g.gather_chan = (int)c->chans.size();
      c->chans.push_back(ch);
    }
  }
  if ((st = invert_coarse(c, g))) return st;
  g.ready = true;
  return FDAL_OK;
}

// ------------------------------------------------------------------ peer channels (setup)
static bool nccl_ok(ncclResult_t r) { return r == ncclSuccess; }
// all-gather `words` int64 per rank through NCCL (setup only)
static int gather_table(fdal_ctx *c, const std::vector<long long> &mine, std::vector<long long> &all) {
  const size_t W = mine.size();
  long long *d = nullptr;
  CU(cudaMalloc((void **)&d, W * (size_t)c->nranks * sizeof(long long)));
  CU(cudaMemcpyAsync(d + W * (size_t)c->rank, mine.data(), W * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  if (!nccl_ok(nccl_api()->AllGather(d + W * (size_t)c->rank, d, W, ncclInt64, c->comm, c->stream))) {
    cudaFree(d);
    set_err(c, "ncclAllGather failed during channel setup");
    return FDAL_ERR_NCCL;
  }
  all.resize(W * (size_t)c->nranks);
  CU(cudaMemcpyAsync(all.data(), d, all.size() * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaFree(d);
  return FDAL_OK;
}
// can every rank map every other rank's memory?  (decided once, before anything is uploaded;
// all ranks take the same decision)
static int probe_peer_access(fdal_ctx *c, bool *ok_out) {
  *ok_out = false;
  char *buf = nullptr;
  CU(cudaMalloc((void **)&buf, 4096));
  cudaIpcMemHandle_t h;
  bool ok = cudaIpcGetMemHandle(&h, buf) == cudaSuccess;
  std::vector<long long> mine(9, 0), all;
  memcpy(mine.data(), &h, sizeof(h));
  mine[8] = ok ? 1 : 0;
  int st = gather_table(c, mine, all);
  if (st) {
    cudaFree(buf);
    return st;
  }
  std::vector<void *> opened;
  for (int q = 0; q < c->nranks && ok; ++q) {
    if (!all[(size_t)q * 9 + 8]) ok = false;
    if (q == c->rank || !ok) continue;
    cudaIpcMemHandle_t hq;
    memcpy(&hq, &all[(size_t)q * 9], sizeof(hq));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      ok = false;
      cudaGetLastError();
    } else {
      opened.push_back(p);
    }
  }
  for (void *p : opened) cudaIpcCloseMemHandle(p);
  // agreement: one more round
  mine.assign(9, 0);
  mine[8] = ok ? 1 : 0;
  st = gather_table(c, mine, all);
  cudaFree(buf);
  if (st) return st;
  for (int q = 0; q < c->nranks; ++q) ok = ok && all[(size_t)q * 9 + 8];
  *ok_out = ok;
  return FDAL_OK;
}
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// allocate the arena, exchange layouts + IPC handles, map the peers, build the device views
static int comm_build(fdal_ctx *c) {
  if (!c->p2p) return FDAL_OK;
  const int nr = c->nranks;
  const size_t nch = c->chans.size();
  size_t off = 0;
  for (Chan &ch : c->chans) {
    ch.recv_off = off;
    off += align256((size_t)2 * std::max(ch.cap, 1) * sizeof(double));
    ch.flags_off = off;
    off += align256((size_t)nr * sizeof(unsigned long long));
    ch.misc_off = off;
    off += 256;  // epoch (8 B) at +0, push counter (4 B) at +128
  }
  c->arena_bytes = std::max<size_t>(off, 256);
  CU(cudaMalloc((void **)&c->arena, c->arena_bytes));
  CU(cudaMemsetAsync(c->arena, 0, c->arena_bytes, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->arena));
  const size_t per = 3 + (size_t)nr, W = 8 + 1 + nch * per;
  std::vector<long long> mine(W, 0), all;
  memcpy(mine.data(), &h, sizeof(h));
  mine[8] = (long long)nch;
  for (size_t i = 0; i < nch; ++i) {
    const Chan &ch = c->chans[i];
    long long *t = &mine[9 + i * per];
    t[0] = (long long)ch.recv_off;
    t[1] = (long long)ch.flags_off;
    t[2] = ch.cap;
    for (int q = 0; q < nr; ++q) t[3 + q] = ch.src_off[(size_t)q];
  }
  // the channel count must agree before tables of that size are exchanged
  {
    std::vector<long long> cnt(1, (long long)nch), cnts;
    int st = gather_table(c, cnt, cnts);
    if (st) return st;
    for (int q = 0; q < nr; ++q)
      if (cnts[(size_t)q] != (long long)nch) {
        set_err(c, "rank %d registered %lld exchange channels, rank %d has %zu: the ranks were set up differently", q,
                cnts[(size_t)q], c->rank, nch);
        return FDAL_ERR_STATE;
      }
  }
  int st = gather_table(c, mine, all);
  if (st) return st;
  c->peer_arena.assign((size_t)nr, nullptr);
  c->peer_arena[(size_t)c->rank] = c->arena;
  for (int q = 0; q < nr; ++q) {
    if (q == c->rank) continue;
    cudaIpcMemHandle_t hq;
    memcpy(&hq, &all[(size_t)q * W], sizeof(hq));
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hq, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_err(c, "cudaIpcOpenMemHandle(rank %d) failed: %s (set FDAL_COMM=nccl for the NCCL fallback)", q,
              cudaGetErrorString(e));
      return FDAL_ERR_CUDA;
    }
    c->peer_arena[(size_t)q] = (char *)p;
  }
  auto T = [&](int q, size_t i) { return &all[(size_t)q * W + 9 + i * per]; };
  // one device blob with every channel's neighbour tables
  std::vector<char> blob;
  auto put = [&](const void *src, size_t bytes) {
    const size_t o = align256(blob.size());
    blob.resize(o + bytes);
    if (bytes) memcpy(blob.data() + o, src, bytes);
    return o;
  };
  struct Offs {
    size_t rank, recv, cap, flag, begin;
  };
  std::vector<Offs> offs(nch);
  for (size_t i = 0; i < nch; ++i) {
    Chan &ch = c->chans[i];
    const int nnb = (int)ch.nb.size();
    std::vector<double *> nrecv((size_t)nnb);
    std::vector<int> ncap((size_t)nnb);
    std::vector<unsigned long long *> nflag((size_t)nnb);
    for (int k = 0; k < nnb; ++k) {
      const int q = ch.nb[(size_t)k];
      const long long *t = T(q, i);
      // where MY entries start inside q's slot
      const long long o = ch.bcast ? ch.src_off[(size_t)c->rank] : t[3 + c->rank];
      if (ch.bcast && (t[2] != ch.cap)) {
        set_err(c, "exchange channel %zu: rank %d has capacity %lld, rank %d has %d", i, q, t[2], c->rank, ch.cap);
        return FDAL_ERR_SHAPE;
      }
      nrecv[(size_t)k] = (double *)(c->peer_arena[(size_t)q] + t[0]) + std::max<long long>(o, 0);
      ncap[(size_t)k] = (int)t[2];
      nflag[(size_t)k] = (unsigned long long *)(c->peer_arena[(size_t)q] + t[1]) + c->rank;
    }
    if (ch.nb_begin.empty()) ch.nb_begin.assign((size_t)nnb + 1, 0);
    offs[i].rank = put(ch.nb.data(), (size_t)nnb * sizeof(int));
    offs[i].recv = put(nrecv.data(), (size_t)nnb * sizeof(double *));
    offs[i].cap = put(ncap.data(), (size_t)nnb * sizeof(int));
    offs[i].flag = put(nflag.data(), (size_t)nnb * sizeof(unsigned long long *));
    offs[i].begin = put(ch.nb_begin.data(), ch.nb_begin.size() * sizeof(int));
  }
  const size_t views_off = align256(blob.size());
  blob.resize(views_off + nch * sizeof(ChanDev));
  char *dblob = nullptr;
  if ((st = dmalloc(c, &dblob, blob.size()))) return st;
  for (size_t i = 0; i < nch; ++i) {
    Chan &ch = c->chans[i];
    ch.d.recv = (double *)(c->arena + ch.recv_off);
    ch.d.flags = (unsigned long long *)(c->arena + ch.flags_off);
    ch.d.epoch = (unsigned long long *)(c->arena + ch.misc_off);
    ch.d.counter = (unsigned int *)(c->arena + ch.misc_off + 128);
    ch.d.cap = ch.cap;
    ch.d.nnb = (int)ch.nb.size();
    ch.d.nb_rank = (const int *)(dblob + offs[i].rank);
    ch.d.nb_recv = (double *const *)(dblob + offs[i].recv);
    ch.d.nb_cap = (const int *)(dblob + offs[i].cap);
    ch.d.nb_flag = (unsigned long long *const *)(dblob + offs[i].flag);
    ch.d.nb_begin = (const int *)(dblob + offs[i].begin);
    ch.d_dev = (ChanDev *)(dblob + views_off) + i;
    memcpy(blob.data() + views_off + i * sizeof(ChanDev), &ch.d, sizeof(ChanDev));
  }
  CU(cudaMemcpyAsync(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  // nobody may push before every rank has zeroed and mapped everything: one more collective
  {
    std::vector<long long> one(1, 1), ones;
    if ((st = gather_table(c, one, ones))) return st;
  }
  return FDAL_OK;
}

static int alloc_cg(fdal_ctx *c, CgWs &w, int64_t n) {
  int st;
  w.n = n;
  if ((st = dvec(c, &w.r, n))) return st;
  if ((st = dvec(c, &w.z, n))) return st;
  if ((st = dvec(c, &w.p, n))) return st;
  if ((st = dvec(c, &w.v, n))) return st;
  if ((st = dvec(c, &w.x, n))) return st;
  if ((st = dvec(c, &w.bin, n))) return st;
  if ((st = dvec(c, &w.scal, S_COUNT))) return st;
  return FDAL_OK;
}
static int need(fdal_ctx *c, int id, const char *name) {
  if (!c->hmat[id].set) {
    set_err(c, "matrix %s not set", name);
    return 0;
  }
  return 1;
}
static int invdiag_of(fdal_ctx *c, const DevCsr &M, double **out) {
  int st = dvec(c, out, M.d.nrows);
  if (st) return st;
  k_inv_diag_from_csr<<<std::max(1, (M.d.nrows + 255) / 256), 256, 0, c->stream>>>(M.d.nrows, M.rp, M.ci, M.v, *out);
  return FDAL_OK;
}

}  // namespace fdal

// ====================================================================== C ABI
#define CHECK_CTX(c) \
  if (!(c)) return FDAL_ERR_INVALID
#define NEED_FINAL(c)                              \
  if (!(c)->finalized) {                           \
    set_err((c), "call fdal_finalize first");      \
    return FDAL_ERR_STATE;                         \
  }

extern "C" {

const char *fdal_version(void) { return "fdal 0.1 (sm_100a CUDA, FP64)"; }

int fdal_create(fdal_ctx **out, const fdal_config *cfg) {
  if (!out || !cfg) return FDAL_ERR_INVALID;
  if (cfg->kind < 0 || cfg->kind > FDAL_KIND_ELLIPTIC_MODIFIED) return FDAL_ERR_INVALID;
  if (cfg->restart > 120) return FDAL_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return FDAL_ERR_CUDA;  // no CPU fallback
  if (cfg->device < 0 || cfg->device >= ndev) return FDAL_ERR_INVALID;
  fdal_ctx *c = new (std::nothrow) fdal_ctx();
  if (!c) return FDAL_ERR_ALLOC;
  c->cfg = *cfg;
  if (c->cfg.restart <= 0) c->cfg.restart = 30;
  if (cudaSetDevice(cfg->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return FDAL_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, cfg->device);
  if (cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return FDAL_ERR_CUDA;
  }
  if (const char *e = getenv("FDAL_OVERLAP")) c->overlap_mass = atoi(e) != 0;
  // tuning knobs (measurement only; defaults are the shipped configuration)
  if (const char *e = getenv("FDAL_UNROLL")) c->spmv_unroll = atoi(e);
  if (const char *e = getenv("FDAL_BSR_UNROLL")) c->bsr_unroll2 = c->bsr_unroll3 = atoi(e);
  *out = c;
  return FDAL_OK;
}

void fdal_destroy(fdal_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (CgWs *w : {&c->cg11, &c->cg22, &c->cgblk, &c->cgmass_m, &c->cgmass_p}) {
    if (w->body_exec) cudaGraphExecDestroy(w->body_exec);
    if (w->fixed_exec) cudaGraphExecDestroy(w->fixed_exec);
  }
  for (MassCheb *mc : {&c->mcheb_m, &c->mcheb_p})
    if (mc->exec) cudaGraphExecDestroy(mc->exec);
  for (void *p : c->allocs) cudaFree(p);
  for (int q = 0; q < (int)c->peer_arena.size(); ++q)
    if (q != c->rank && c->peer_arena[(size_t)q]) cudaIpcCloseMemHandle(c->peer_arena[(size_t)q]);
  if (c->arena) cudaFree(c->arena);
  if (c->comm) nccl_api()->CommDestroy(c->comm);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *fdal_last_error(const fdal_ctx *c) { return c ? c->err.c_str() : "null context"; }

#define NOT_FINAL(c)                                                                                            \
  if ((c)->finalized || !(c)->allocs.empty()) {                                                                 \
    set_err((c), "the context is already finalized: setters are refused (destroy it and create a new one)");    \
    return FDAL_ERR_STATE;                                                                                      \
  }
// row_ptr / column checks shared by every CSR entry point
static int check_csr(fdal_ctx *c, const char *what, int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp,
                     const int32_t *ci, const double *v) {
  if (!rp || nr < 0 || nc < 0 || nnz < 0 || (nnz && (!ci || !v))) {
    set_err(c, "%s: bad argument", what);
    return FDAL_ERR_INVALID;
  }
  if (rp[0] != 0 || rp[nr] != nnz) {
    set_err(c, "%s: row_ptr inconsistent with nnz", what);
    return FDAL_ERR_SHAPE;
  }
  if (nnz >= (int64_t)std::numeric_limits<int>::max() || nr >= (int64_t)std::numeric_limits<int>::max()) {
    set_err(c, "%s: %lld nnz / %lld rows exceed the 32-bit row_ptr of this build", what, (long long)nnz, (long long)nr);
    return FDAL_ERR_UNSUPPORTED;
  }
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t i = 0; i < nr; ++i) {
    if (rp[i + 1] < rp[i]) {
      bad |= 1;
      continue;
    }
    if (rp[i] < 0 || rp[i + 1] > nnz) {
      bad |= 1;
      continue;
    }
    for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
      if (ci[k] < 0 || ci[k] >= nc) bad |= 2;
  }
  if (bad & 1) {
    set_err(c, "%s: row_ptr not monotone", what);
    return FDAL_ERR_SHAPE;
  }
  if (bad & 2) {
    set_err(c, "%s: column index out of range", what);
    return FDAL_ERR_SHAPE;
  }
  return FDAL_OK;
}
int fdal_set_csr(fdal_ctx *c, int id, int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp, const int32_t *ci,
                 const double *v) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (id < 0 || id >= FDAL_MAT_COUNT) {
    set_err(c, "fdal_set_csr: bad matrix id %d", id);
    return FDAL_ERR_INVALID;
  }
  char what[32];
  snprintf(what, sizeof(what), "matrix %d", id);
  int st = check_csr(c, what, nr, nc, nnz, rp, ci, v);
  if (st) return st;
  fill_host_csr(c->hmat[id], nr, nc, nnz, rp, ci, v);
  return FDAL_OK;
}

int fdal_set_diag(fdal_ctx *c, int id, int64_t n, const double *d) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (!d || n < 0) return FDAL_ERR_INVALID;
  if (id == FDAL_DIAG_W_INV)
    c->h_winv.assign(d, d + n);
  else if (id == FDAL_DIAG_MP_LUMPED_INV)
    c->h_mp_lumped.assign(d, d + n);
  else
    return FDAL_ERR_INVALID;
  return FDAL_OK;
}

static int copy_view(fdal_ctx *c, const char *what, const fdal_csr_view *v, HostCsr &h) {
  int st = check_csr(c, what, v->n_rows, v->n_cols, v->nnz, v->row_ptr, v->col, v->val);
  if (st) return st;
  fill_host_csr(h, v->n_rows, v->n_cols, v->nnz, v->row_ptr, v->col, v->val);
  return FDAL_OK;
}
int fdal_amg_set_level(fdal_ctx *c, int which, int level, const fdal_csr_view *A, const fdal_csr_view *P,
                       const fdal_csr_view *R, const double *inv_diag, double lmax, int degree, double ratio) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (which < 0 || which > 1 || level < 0 || level > 30 || !A) return FDAL_ERR_INVALID;
  Amg &g = c->amg[which];
  if ((int)g.lev.size() < level + 1) g.lev.resize(level + 1);
  AmgLevel &L = g.lev[level];
  char what[48];
  int st;
  snprintf(what, sizeof(what), "AMG %d level %d A", which, level);
  if ((st = copy_view(c, what, A, L.hA))) return st;
  L.hP = HostCsr();
  L.hR = HostCsr();
  if (P) {
    if (P->n_rows != A->n_rows) {
      set_err(c, "AMG level %d: P has %lld rows, A has %lld", level, (long long)P->n_rows, (long long)A->n_rows);
      return FDAL_ERR_SHAPE;
    }
    snprintf(what, sizeof(what), "AMG %d level %d P", which, level);
    if ((st = copy_view(c, what, P, L.hP))) return st;
  }
  if (R) {
    snprintf(what, sizeof(what), "AMG %d level %d R", which, level);
    if ((st = copy_view(c, what, R, L.hR))) return st;
  }
  L.h_invd.clear();
  if (inv_diag) L.h_invd.assign(inv_diag, inv_diag + A->n_rows);
  L.lmax = lmax;
  L.degree = degree;
  L.ratio = ratio;
  return FDAL_OK;
}
int fdal_amg_set_coarse(fdal_ctx *c, int which, int level, const fdal_csr_view *A) {
  return fdal_amg_set_level(c, which, level, A, nullptr, nullptr, nullptr, 1.0, 0, 1.0);
}

int fdal_finalize(fdal_ctx *c) {
  CHECK_CTX(c);
  if (c->finalized) return FDAL_OK;
  if (!c->allocs.empty()) {
    set_err(c, "fdal_finalize may only be called once per context");
    return FDAL_ERR_STATE;
  }
  CU(cudaSetDevice(c->cfg.device));
  const int k = c->cfg.kind;
  int st;
  if (!need(c, FDAL_MAT_A, "A") || !need(c, FDAL_MAT_CT, "Ct")) return FDAL_ERR_STATE;
  c->n0 = c->hmat[FDAL_MAT_A].nr;
  c->m = c->hmat[FDAL_MAT_CT].nc;
  if (c->hmat[FDAL_MAT_CT].nr != c->n0 || c->hmat[FDAL_MAT_A].owned_cols() != c->n0) {
    set_err(c, "A must be n x n and Ct n x m");
    return FDAL_ERR_SHAPE;
  }
  const bool D = c->nranks > 1;
  if (D && is_stokes(c) && !c->hmat[FDAL_MAT_B].set) {
    set_err(c, "a partitioned Stokes system needs an explicit B with its halo plan");
    return FDAL_ERR_STATE;
  }
  if (k == FDAL_KIND_LAPLACE) {
    c->nblocks = 2;
    c->n1 = c->m;
    c->n2 = 0;
  } else if (is_stokes(c)) {
    if (!need(c, FDAL_MAT_BT, "Bt")) return FDAL_ERR_STATE;
    if (c->hmat[FDAL_MAT_BT].nr != c->n0) {
      set_err(c, "Bt must have n_u rows");
      return FDAL_ERR_SHAPE;
    }
    c->nblocks = 3;
    c->n1 = c->hmat[FDAL_MAT_BT].owned_cols();
    c->n2 = c->m;
    if (!need(c, FDAL_MAT_MP, "Mp")) return FDAL_ERR_STATE;
    if (c->hmat[FDAL_MAT_MP].nr != c->n1) return FDAL_ERR_SHAPE;
  } else {
    if (!need(c, FDAL_MAT_A2, "A2") || !need(c, FDAL_MAT_M, "M")) return FDAL_ERR_STATE;
    if (c->hmat[FDAL_MAT_A2].nr != c->m || c->hmat[FDAL_MAT_M].nr != c->m) {
      set_err(c, "A2 and M must be m x m");
      return FDAL_ERR_SHAPE;
    }
    c->nblocks = 3;
    c->n1 = c->m;
    c->n2 = c->m;
  }
  c->N = c->n0 + c->n1 + c->n2;
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    if ((int64_t)c->h_winv.size() != c->m) {
      set_err(c, "diagonal W^-1 of size m=%lld required", (long long)c->m);
      return FDAL_ERR_STATE;
    }
  } else if (!need(c, FDAL_MAT_M, "M (exact W^-1)"))
    return FDAL_ERR_STATE;

  // reductions + scalars
  if ((st = dmalloc(c, &c->d_partials, (size_t)kMaxPartials * 128))) return st;
  if ((st = dmalloc(c, &c->d_counter, 4))) return st;
  CU(cudaMemsetAsync(c->d_counter, 0, 4 * sizeof(unsigned int), c->stream));
  if ((st = dvec(c, &c->d_scal, 512))) return st;
  CU(cudaMallocHost((void **)&c->h_scal, 512 * sizeof(double)));

  if (c->p2p) {
    c->vec_cap = (int)std::max<int64_t>(c->m, 1);
    for (int which = 0; which < 2; ++which) {
      Chan ch;
      ch.bcast = true;
      const int region = which == 0 ? kArCap : c->vec_cap;
      ch.cap = region * c->nranks;
      ch.src_off.assign((size_t)c->nranks, -1);
      ch.src_off[(size_t)c->rank] = (int64_t)region * c->rank;
      for (int q = 0; q < c->nranks; ++q) ch.nb.push_back(q);
      (which == 0 ? c->ch_scalar : c->ch_vec) = (int)c->chans.size();
      c->chans.push_back(ch);
    }
  }
  PhaseTimer pt;
  // explicit transposes (gather kernels only: no atomics, deterministic)
  if (!c->hmat[FDAL_MAT_C].set) host_transpose(c->hmat[FDAL_MAT_CT], c->hmat[FDAL_MAT_C]);
  if (is_stokes(c) && !c->hmat[FDAL_MAT_B].set) host_transpose(c->hmat[FDAL_MAT_BT], c->hmat[FDAL_MAT_B]);
  pt.mark("transposes C = Ct^T, B = Bt^T (host)");
  for (int id = 0; id < FDAL_MAT_COUNT; ++id)
    if (c->hmat[id].set)
      if ((st = upload_csr(c, c->hmat[id], c->dmat[id], id == FDAL_MAT_A ? c->cfg.block_size : 1))) return st;
  pt.mark("system matrices: upload (+ BSR of A)", c->hmat[FDAL_MAT_A].nnz);
  // Ct rides along in the row pass over A (fused augmented apply): flag the row chunks of A in which Ct
  // has any entry, so the other ~99 % of the chunks never look at Ct's row pointers
  {
    const DevCsr &A = c->dmat[FDAL_MAT_A];
    const HostCsr &hct = c->hmat[FDAL_MAT_CT];
    const int64_t rpb = (int64_t)(kBlock / (A.use_bsr ? A.bsr.tpr : A.d.tpr)) * (A.use_bsr ? A.bsr_b : 1);
    const int64_t nchunks = (hct.nr + rpb - 1) / rpb;
    std::vector<unsigned char> any((size_t)std::max<int64_t>(nchunks, 1), 0);
    for (int64_t q = 0; q < nchunks; ++q)
      any[(size_t)q] = hct.rp[(size_t)std::min<int64_t>(hct.nr, (q + 1) * rpb)] > hct.rp[(size_t)(q * rpb)];
    unsigned char *d_any = nullptr;
    if ((st = dmalloc(c, &d_any, any.size()))) return st;
    CU(cudaMemcpyAsync(d_any, any.data(), any.size(), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!getenv("FDAL_NO_CHUNK_FLAGS")) c->dmat[FDAL_MAT_CT].d.chunk_any = d_any;
  }
  for (int id = 0; id < FDAL_MAT_COUNT; ++id) {  // host copies no longer needed
    HostCsr &h = c->hmat[id];
    IntBuf().swap(h.ci);
    DblBuf().swap(h.v);
    IntBuf().swap(h.rp);
  }
  for (int id : {FDAL_MAT_A, FDAL_MAT_BT, FDAL_MAT_B, FDAL_MAT_MP}) c->dmat[id].dist_rows = D;
  // AMG
  if (c->cfg.inner_prec == FDAL_PREC_AMG) {
    for (int a = 0; a < 2; ++a) {
      if (a == 1 && !is_elliptic(c)) continue;
      if (c->amg[a].lev.empty()) {
        set_err(c, "AMG hierarchy %d not set", a);
        return FDAL_ERR_STATE;
      }
      if (c->amg[a].lev.size() > 1 && c->amg[a].lev[0].hA.nr != (a == 0 ? c->n0 : c->n1)) {
        set_err(c, "AMG hierarchy %d: fine level has %lld rows, block has %lld", a,
                (long long)c->amg[a].lev[0].hA.nr, (long long)(a == 0 ? c->n0 : c->n1));
        return FDAL_ERR_SHAPE;
      }
      if ((st = prepare_amg(c, c->amg[a], a == 0 && c->nranks > 1))) return st;
    }
  }
  pt.mark("AMG hierarchies: upload (+ BSR of level 0)");
  // peer channels: everything that exchanges data has been registered by now
  if ((st = comm_build(c))) return st;
  // dots: the replicated tail blocks are counted on rank 0 only
  {
    const int64_t tail = k == FDAL_KIND_LAPLACE ? c->n1 : is_stokes(c) ? c->n2 : c->n1 + c->n2;
    c->n_dot_outer = c->N - ((D && c->rank != 0) ? tail : 0);
  }
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    if ((st = dvec(c, &c->d_winv, c->m))) return st;
    CU(cudaMemcpyAsync(c->d_winv, c->h_winv.data(), (size_t)c->m * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  // scratch
  if ((st = dvec(c, &c->t_m0, c->m))) return st;
  if ((st = dvec(c, &c->t_m1, c->m))) return st;
  if ((st = dvec(c, &c->t_m2, c->m))) return st;
  if ((st = dvec(c, &c->t_mw, c->m))) return st;
  if ((st = dvec(c, &c->t_n0, c->n0))) return st;
  if ((st = dvec(c, &c->t_p0, c->n1))) return st;
  if ((st = dvec(c, &c->t_p1, c->n1))) return st;
  if ((st = dvec(c, &c->t_N0, c->N))) return st;
  if ((st = dvec(c, &c->t_N1, c->N))) return st;
  if ((st = dvec(c, &c->t_N2, c->N))) return st;
  if ((st = alloc_cg(c, c->cg11, c->n0))) return st;
  c->cg11.dist = D;
  c->cg11.n_dot = c->n0;
  if (is_elliptic(c)) {
    if ((st = alloc_cg(c, c->cg22, c->n1))) return st;  // immersed block: replicated
    if (k == FDAL_KIND_ELLIPTIC_IDEAL) {
      if ((st = alloc_cg(c, c->cgblk, c->n0 + c->n1))) return st;
      c->cgblk.dist = D;
      c->cgblk.n_dot = c->n0 + ((D && c->rank != 0) ? 0 : c->n1);
    }
  }
  // mass solves
  if (c->cfg.winv_mode != FDAL_WINV_DIAG) {
    if ((st = alloc_cg(c, c->cgmass_m, c->m))) return st;
    if ((st = invdiag_of(c, c->dmat[FDAL_MAT_M], &c->d_m_invdiag))) return st;
    CgHistory hist_m;
    if ((st = mass_calibrate(c, c->cgmass_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, &c->mass_its_m, &hist_m))) return st;
    // multiplier spaces too large for a dense inverse: Chebyshev form (elliptic_interface, co-dimension 0).  The
    // single-CTA Jacobi-PCG that used to cover 4096 < m <= 16384 is 3x slower (cycle 5, m = 5 249: 2994 -> 954 ms
    // per solve) and stays only as the FDAL_MASS_CHEB=0 fallback
    // (FDAL_MASS_CHEB_MIN_ROWS: test knob, lets small problems take this path)
    const char *cmin = getenv("FDAL_MASS_CHEB_MIN_ROWS");
    if (c->m > (cmin ? atoll(cmin) : (long long)kDenseWinvMaxRows) &&
        (st = mass_cheb_setup(c, c->mcheb_m, c->cgmass_m, c->dmat[FDAL_MAT_M], c->d_m_invdiag, hist_m)))
      return st;
    static const bool no_cta = getenv("FDAL_NO_MASS_CTA") != nullptr;
    if (c->m <= kMassCtaMaxRows && !no_cta && c->mcheb_m.mode == 0) {
      if ((st = dvec(c, &c->mass_cta_ws, 5 * c->m))) return st;
      CU(cudaFuncSetAttribute(k_mass_pcg_cta<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(4 * kMassCtaSmemRows * sizeof(double))));
    }
    // small multiplier spaces: the exact W^-1 as ONE L2-resident dense GEMV (FDAL_DENSE_WINV=0: keep the PCG)
    const char *dw = getenv("FDAL_DENSE_WINV");
    if ((!dw || atoi(dw) > 0) && c->m > 0 && c->m <= kDenseWinvMaxRows && c->mcheb_m.mode == 0 &&
        (st = build_dense_winv(c)))
      return st;
  }
  if (is_stokes(c)) {
    if ((st = alloc_cg(c, c->cgmass_p, c->n1))) return st;
    c->cgmass_p.dist = D;
    c->cgmass_p.n_dot = c->n1;
    if (c->cfg.mp_inv_mode == FDAL_MPINV_EXACT) {
      if ((st = invdiag_of(c, c->dmat[FDAL_MAT_MP], &c->d_mp_invdiag))) return st;
      CgHistory hist_p;
      if ((st = mass_calibrate(c, c->cgmass_p, c->dmat[FDAL_MAT_MP], c->d_mp_invdiag, &c->mass_its_p, &hist_p)))
        return st;
      if ((st = mass_cheb_setup(c, c->mcheb_p, c->cgmass_p, c->dmat[FDAL_MAT_MP], c->d_mp_invdiag, hist_p))) return st;
    } else {
      if ((st = dvec(c, &c->d_mp_lumped, c->n1))) return st;
      if ((int64_t)c->h_mp_lumped.size() == c->n1) {
        CU(cudaMemcpyAsync(c->d_mp_lumped, c->h_mp_lumped.data(), (size_t)c->n1 * sizeof(double),
                           cudaMemcpyHostToDevice, c->stream));
      } else {
        // M_p * 1, inverted (stokes_immersed_boundary.cc:946-952)
        k_fill<<<grid_elems(c, c->n1), kBlock, 0, c->stream>>>(c->n1, c->t_p0, 1.0);
        spmv(c, c->dmat[FDAL_MAT_MP], c->t_p0, EpiAssign{c->d_mp_lumped, 1.0});
        k_reciprocal<<<grid_elems(c, c->n1), kBlock, 0, c->stream>>>(c->n1, c->d_mp_lumped);
      }
    }
  }
  // outer Krylov workspace
  const int mb = c->cfg.restart;
  if (k == FDAL_KIND_STOKES_DIAG_MINRES) {
    for (int i = 0; i < 3; ++i) {
      if ((st = dvec(c, &c->mr_u[i], c->N))) return st;
      if ((st = dvec(c, &c->mr_m[i], c->N))) return st;
    }
    if ((st = dvec(c, &c->mr_v, c->N))) return st;
  } else {
    if ((st = dvec(c, &c->V, (int64_t)(mb + 1) * c->N))) return st;
    if ((st = dvec(c, &c->Z, (int64_t)mb * c->N))) return st;
  }
  if ((st = dvec(c, &c->d_h, 2 * (mb + 2) + 8))) return st;
  if ((st = dvec(c, &c->d_y, mb + 8))) return st;
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  pt.mark("workspaces, mass-solve calibration, Krylov bases");
  c->finalized = true;
  return FDAL_OK;
}

int fdal_block_sizes(const fdal_ctx *c, int64_t sizes[3], int *nb) {
  CHECK_CTX(c);
  sizes[0] = c->n0;
  sizes[1] = c->n1;
  sizes[2] = c->n2;
  *nb = c->nblocks;
  return FDAL_OK;
}

// ---- host-pointer wrappers: stage through t_N0 / t_N1 -----------------------------
static int h2d(fdal_ctx *c, double *d, const double *h, int64_t n) {
  CU(cudaMemcpyAsync(d, h, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return FDAL_OK;
}
static int d2h(fdal_ctx *c, double *h, const double *d, int64_t n) {
  CU(cudaMemcpyAsync(h, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaGetLastError());
  return FDAL_OK;
}
#define BEGIN_CALL(c)              \
  CHECK_CTX(c);                    \
  NEED_FINAL(c);                   \
  CU(cudaSetDevice((c)->cfg.device)); \
  (c)->fail = 0

int fdal_spmv_dev(fdal_ctx *c, int id, int transpose, const double *d_x, double *d_y) {
  BEGIN_CALL(c);
  if (id < 0 || id >= FDAL_MAT_COUNT) return FDAL_ERR_INVALID;
  int use = id;
  if (transpose) {
    if (id == FDAL_MAT_CT) use = FDAL_MAT_C;
    else if (id == FDAL_MAT_C) use = FDAL_MAT_CT;
    else if (id == FDAL_MAT_BT) use = FDAL_MAT_B;
    else if (id == FDAL_MAT_B) use = FDAL_MAT_BT;
    else if (id == FDAL_MAT_A || id == FDAL_MAT_A2 || id == FDAL_MAT_M || id == FDAL_MAT_MP) use = id;  // symmetric
  }
  if (!c->dmat[use].set) {
    set_err(c, "matrix %d not available", use);
    return FDAL_ERR_INVALID;
  }
  spmv(c, c->dmat[use], d_x, EpiAssign{d_y, 1.0});
  CU(cudaStreamSynchronize(c->stream));
  return FDAL_OK;
}
int fdal_spmv(fdal_ctx *c, int id, int transpose, const double *x, double *y) {
  BEGIN_CALL(c);
  if (id < 0 || id >= FDAL_MAT_COUNT || !c->dmat[id].set) return FDAL_ERR_INVALID;
  const int64_t nin = transpose ? c->dmat[id].d.nrows : c->dmat[id].d.ncols;
  const int64_t nout = transpose ? c->dmat[id].d.ncols : c->dmat[id].d.nrows;
  int st;
  if ((st = h2d(c, c->t_N0, x, nin))) return st;
  if ((st = fdal_spmv_dev(c, id, transpose, c->t_N0, c->t_N1))) return st;
  return d2h(c, y, c->t_N1, nout);
}
int fdal_apply_aug_dev(fdal_ctx *c, int which, const double *d_x, double *d_y) {
  BEGIN_CALL(c);
  if (which == FDAL_AMG_A22 && !is_elliptic(c)) return FDAL_ERR_INVALID;
  apply_aug(c, which, d_x, d_y, nullptr);
  CU(cudaStreamSynchronize(c->stream));
  return c->fail;
}
int fdal_apply_aug(fdal_ctx *c, int which, const double *x, double *y) {
  BEGIN_CALL(c);
  const int64_t n = which == FDAL_AMG_A11 ? c->n0 : c->n1;
  int st;
  if ((st = h2d(c, c->t_N0, x, n))) return st;
  if ((st = fdal_apply_aug_dev(c, which, c->t_N0, c->t_N1))) return st;
  return d2h(c, y, c->t_N1, n);
}
int fdal_apply_system(fdal_ctx *c, const double *x, double *y) {
  BEGIN_CALL(c);
  int st;
  if ((st = h2d(c, c->t_N0, x, c->N))) return st;
  apply_system(c, c->t_N0, c->t_N1);
  if ((st = d2h(c, y, c->t_N1, c->N))) return st;
  return c->fail;
}
int fdal_apply_winv(fdal_ctx *c, const double *x, double *y) {
  BEGIN_CALL(c);
  int st;
  if ((st = h2d(c, c->t_N0, x, c->m))) return st;
  apply_winv_scaled(c, 1.0, c->t_N0, c->t_N1);
  return d2h(c, y, c->t_N1, c->m);
}
int fdal_apply_mp_inv(fdal_ctx *c, const double *x, double *y, int *its) {
  BEGIN_CALL(c);
  if (!is_stokes(c)) return FDAL_ERR_INVALID;
  int st;
  c->its_mass = 0;
  if ((st = h2d(c, c->t_N0, x, c->n1))) return st;
  st = apply_mp_inv(c, c->t_N0, c->t_N1);
  if (its) *its = c->its_mass;
  if (st) return st;
  return d2h(c, y, c->t_N1, c->n1);
}
int fdal_apply_amg_dev(fdal_ctx *c, int which, const double *d_r, double *d_z) {
  BEGIN_CALL(c);
  if (which < 0 || which > 1 || !c->amg[which].ready) return FDAL_ERR_STATE;
  vcycle(c, c->amg[which], d_r, d_z, nullptr);
  CU(cudaStreamSynchronize(c->stream));
  return FDAL_OK;
}
int fdal_apply_amg(fdal_ctx *c, int which, const double *r, double *z) {
  BEGIN_CALL(c);
  if (which < 0 || which > 1 || !c->amg[which].ready) return FDAL_ERR_STATE;
  const int64_t n = c->amg[which].lev[0].n;
  int st;
  if ((st = h2d(c, c->t_N0, r, n))) return st;
  if ((st = fdal_apply_amg_dev(c, which, c->t_N0, c->t_N1))) return st;
  return d2h(c, z, c->t_N1, n);
}
int fdal_apply_aug_inv(fdal_ctx *c, int which, const double *b, double *x, int *its) {
  BEGIN_CALL(c);
  if (which == FDAL_AMG_A22 && !is_elliptic(c)) return FDAL_ERR_INVALID;
  const int64_t n = which == FDAL_AMG_A11 ? c->n0 : c->n1;
  int st, i = 0;
  if ((st = h2d(c, c->t_N0, b, n))) return st;
  st = apply_aug_inv(c, which, c->t_N0, c->t_N1, &i);
  if (its) *its = i;
  if (st) {
    set_err(c, "inner CG did not converge in %d steps (SolverControl::NoConvergence)", i);
    return st;
  }
  return d2h(c, x, c->t_N1, n);
}
int fdal_apply_prec(fdal_ctx *c, const double *u, double *v, int inner_its[2]) {
  BEGIN_CALL(c);
  int st;
  c->its_a11 = c->its_a22 = 0;
  // apply_prec uses t_N0 (ideal variant): stage through V-independent buffers
  double *du = c->t_N1, *dv = c->t_N2;
  if ((st = h2d(c, du, u, c->N))) return st;
  st = apply_prec(c, du, dv);
  if (inner_its) {
    inner_its[0] = c->its_a11;
    inner_its[1] = c->its_a22;
  }
  int st2 = d2h(c, v, dv, c->N);
  if (st) set_err(c, "inner solver did not converge (SolverControl::NoConvergence)");
  return st ? st : st2;
}
int fdal_augment_rhs(fdal_ctx *c, double *rhs) {
  BEGIN_CALL(c);
  int st;
  if ((st = h2d(c, c->t_N0, rhs, c->N))) return st;
  double *g = c->t_N0 + c->n0 + (c->nblocks == 3 ? c->n1 : 0);
  apply_winv_scaled(c, c->cfg.gamma, g, c->t_m0);
  spmv(c, c->dmat[FDAL_MAT_CT], c->t_m0, EpiAdd{c->t_N0, 1.0});
  return d2h(c, rhs, c->t_N0, c->N);
}

int fdal_solve_dev(fdal_ctx *c, const double *d_rhs, double *d_x, fdal_solve_info *info) {
  BEGIN_CALL(c);
  fdal_solve_info local;
  if (!info) info = &local;
  memset(info, 0, sizeof(*info));
  c->its_a11 = c->its_a22 = c->its_mass = c->n_inner_solves = 0;
  const int64_t l0 = c->launches, g0 = c->graph_launches;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, c->stream));
  int st = c->cfg.kind == FDAL_KIND_STOKES_DIAG_MINRES ? minres(c, d_rhs, d_x, info) : fgmres(c, d_rhs, d_x, info);
  cudaEventRecord(e1, c->stream);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  info->solve_ms = ms;
  info->status = st;
  info->inner_iterations = c->its_a11;
  info->inner_iterations_a22 = c->its_a22;
  info->inner_solves = c->n_inner_solves;
  info->mass_iterations = c->its_mass;
  info->kernel_launches = c->launches - l0;
  info->reserved = (int32_t)std::min<int64_t>(c->graph_launches - g0, std::numeric_limits<int32_t>::max());
  if (st == FDAL_ERR_INNER_NO_CONVERGENCE)
    set_err(c, "inner CG did not converge (SolverControl::NoConvergence)");
  else if (st == FDAL_ERR_OUTER_NO_CONVERGENCE)
    set_err(c, "outer solver did not converge after %d steps", info->outer_iterations);
  else if (st == FDAL_ERR_MASS_NO_CONVERGENCE)
    set_err(c, "pressure mass CG did not converge");
  return st;
}
int fdal_solve(fdal_ctx *c, const double *rhs, double *x, fdal_solve_info *info) {
  BEGIN_CALL(c);
  int st;
  double *drhs = c->t_N1, *dx = c->t_N2;
  if ((st = h2d(c, drhs, rhs, c->N))) return st;
  if ((st = h2d(c, dx, x, c->N))) return st;
  st = fdal_solve_dev(c, drhs, dx, info);
  int st2 = d2h(c, x, dx, c->N);
  return st ? st : st2;
}

// ---- measurement ----------------------------------------------------------------------
static double csr_bytes(const CsrDev &A) {
  return 12.0 * (double)A.nnz + 4.0 * ((double)A.nrows + 1) + 8.0 * (double)A.ncols + 8.0 * (double)A.nrows;
}
// algorithmic bytes of one mat-vec with the storage actually used (BSR or CSR)
static double mat_bytes(const DevCsr &A) {
  if (!A.use_bsr) return csr_bytes(A.d);
  const double b2 = (double)A.bsr_b * A.bsr_b;
  return (8.0 * b2 + 4.0) * (double)A.bsr_nblocks + 4.0 * ((double)A.bsr.nbrows + 1) + 8.0 * (double)A.d.ncols +
         8.0 * (double)A.d.nrows;
}
int fdal_time_kernel(fdal_ctx *c, int what, int param, int warmup, int reps, int flush_l2, double *avg_ms,
                     double *alg_bytes, int64_t *launches_per_rep) {
  BEGIN_CALL(c);
  int st;
  if (flush_l2 && !c->flush_buf) {
    c->flush_bytes = (size_t)256 << 20;
    if ((st = dmalloc(c, &c->flush_buf, c->flush_bytes))) return st;
  }
  const DevCsr &A = c->dmat[FDAL_MAT_A];
  double *x = c->t_N0, *y = c->t_N1;
  k_fill<<<grid_elems(c, c->N), kBlock, 0, c->stream>>>(c->N, x, 1.0);
  double bytes = 0;
  std::function<void()> run;
  const int64_t N = c->N, n0 = c->n0;
  switch (what) {
    case FDAL_TIME_SPMV_A:
      run = [&]() { spmv(c, A, x, EpiAssign{y, 1.0}); };
      bytes = mat_bytes(A);
      break;
    case FDAL_TIME_AUG:
      run = [&]() { apply_aug11(c, x, y, nullptr); };
      bytes = mat_bytes(A) + csr_bytes(c->dmat[FDAL_MAT_C].d) + 12.0 * (double)c->dmat[FDAL_MAT_CT].d.nnz +
              4.0 * ((double)n0 + 1) + 16.0 * (double)c->m;
      break;
    case FDAL_TIME_VCYCLE: {
      if (!c->amg[0].ready) return FDAL_ERR_STATE;
      run = [&]() { vcycle(c, c->amg[0], x, y, nullptr); };
      Amg &g = c->amg[0];
      for (size_t l = 0; l + 1 < g.lev.size(); ++l) {
        const AmgLevel &L = g.lev[l];
        const double nl = (double)L.n;
        const int deg = L.degree;
        // pre: zero step (3 vec) + (deg-1) fused steps; residual; R; P; post: deg fused steps
        bytes += 24.0 * nl + (2 * deg - 1) * (mat_bytes(L.A) + 32.0 * nl) + (mat_bytes(L.A) + 8.0 * nl) +
                 csr_bytes(L.R.d) + csr_bytes(L.P.d) + 8.0 * nl;
      }
      const double cn = (double)g.lev.back().n;
      bytes += 8.0 * cn * cn + 16.0 * cn;
    } break;
    case FDAL_TIME_CHEB_FINE: {
      if (!c->amg[0].ready || c->amg[0].lev.size() < 2) return FDAL_ERR_STATE;
      AmgLevel &L = c->amg[0].lev[0];
      run = [&]() {
        EpiCheb<false> e{x, L.invd, L.xa, L.d, L.xb, 0.3, 0.7, 0};
        spmv(c, L.A, L.xa, e);
      };
      bytes = mat_bytes(L.A) + 32.0 * (double)L.n;
    } break;
    case FDAL_TIME_DOT:
      run = [&]() { dot(c, N, x, y, c->d_scal); };
      bytes = 16.0 * (double)N;
      break;
    case FDAL_TIME_MULTIDOT: {
      if (!c->V) return FDAL_ERR_STATE;
      const int nv = std::max(1, std::min(param, c->cfg.restart));
      run = [&, nv]() {
        k_multidot<<<grid_elems(c, N), kBlock, 0, c->stream>>>(N, N, x, c->V, N, nv, 0, reducer(c, c->d_h));
        c->launches++;
      };
      bytes = 8.0 * (double)N * (nv + (nv + kMultiDotGroup - 1) / kMultiDotGroup);
    } break;
    case FDAL_TIME_AXPY:
      run = [&]() { axpby(c, N, 0.5, x, 1.0, y); };
      bytes = 24.0 * (double)N;
      break;
    default:
      return FDAL_ERR_INVALID;
  }
  for (int i = 0; i < warmup; ++i) run();
  CU(cudaStreamSynchronize(c->stream));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  double total = 0;
  const int64_t l0 = c->launches;
  for (int i = 0; i < reps; ++i) {
    if (flush_l2) CU(cudaMemsetAsync(c->flush_buf, i & 0xff, c->flush_bytes, c->stream));
    CU(cudaEventRecord(e0, c->stream));
    run();
    CU(cudaEventRecord(e1, c->stream));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  CU(cudaGetLastError());
  if (avg_ms) *avg_ms = total / std::max(1, reps);
  if (alg_bytes) *alg_bytes = bytes;
  if (launches_per_rep) *launches_per_rep = (c->launches - l0) / std::max(1, reps);
  return FDAL_OK;
}

// ---- multi-GPU entry points ---------------------------------------------------------------
int fdal_nccl_unique_id(char id_out[128]) {
  NcclApi *n = nccl_api();
  if (!n->ok) return FDAL_ERR_NCCL;
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return FDAL_ERR_NCCL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id_out, &id, 128);
  return FDAL_OK;
}
int fdal_comm_init(fdal_ctx *c, const char id[128], int rank, int n_ranks) {
  CHECK_CTX(c);
  if (rank < 0 || n_ranks < 1 || rank >= n_ranks || !id) return FDAL_ERR_INVALID;
  if (c->finalized || !c->allocs.empty()) {
    set_err(c, "fdal_comm_init must precede fdal_finalize");
    return FDAL_ERR_STATE;
  }
  c->rank = rank;
  c->nranks = n_ranks;
  if (n_ranks == 1) return FDAL_OK;
  NcclApi *n = nccl_api();
  if (!n->ok) {
    set_err(c, "libnccl.so.2 could not be loaded");
    return FDAL_ERR_NCCL;
  }
  CU(cudaSetDevice(c->cfg.device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = n->CommInitRank(&c->comm, n_ranks, uid, rank);
  if (r != ncclSuccess) {
    set_err(c, "ncclCommInitRank failed: %s", n->GetErrorString ? n->GetErrorString(r) : "?");
    return FDAL_ERR_NCCL;
  }
  // per-iteration traffic goes through peer-mapped channels unless the ranks cannot map each
  // other's memory (or FDAL_COMM=nccl asks for the NCCL send/recv + all-reduce fallback)
  const char *mode = getenv("FDAL_COMM");
  if (!(mode && strcmp(mode, "nccl") == 0)) {
    bool ok = false;
    int st = probe_peer_access(c, &ok);
    if (st) return st;
    c->p2p = ok;
    if (!ok && mode && strcmp(mode, "p2p") == 0) {
      set_err(c, "FDAL_COMM=p2p but the ranks cannot map each other's memory (cudaIpc)");
      return FDAL_ERR_UNSUPPORTED;
    }
  }
  return FDAL_OK;
}
int fdal_comm_mode(const fdal_ctx *c) { return c ? (c->nranks <= 1 ? 0 : (c->p2p ? 2 : 1)) : -1; }
int fdal_bsr_conversions(const fdal_ctx *c, int32_t *on_device, int32_t *on_host) {
  if (!c || !c->finalized) return FDAL_ERR_STATE;
  if (on_device) *on_device = c->bsr_built_on_device;
  if (on_host) *on_host = c->bsr_built_on_host;
  return FDAL_OK;
}
int fdal_mass_solver_info(const fdal_ctx *c, int which, int32_t *form, int32_t *iterations, double *interval_lo,
                          double *interval_hi, double *verified_residual) {
  if (!c || !c->finalized || which < 0 || which > 1) return FDAL_ERR_STATE;
  const MassCheb &mc = which == 0 ? c->mcheb_m : c->mcheb_p;
  const bool exact = which == 0 ? c->cfg.winv_mode != FDAL_WINV_DIAG
                                : (is_stokes(c) && c->cfg.mp_inv_mode == FDAL_MPINV_EXACT);
  int f = FDAL_MASS_NONE, its = 0;
  if (exact) {
    if (mc.mode == 2)
      f = FDAL_MASS_CHEB_PERSISTENT, its = mc.its;
    else if (mc.mode == 1)
      f = FDAL_MASS_CHEB_KERNELS, its = mc.its;
    else if (which == 0 && c->d_winv_dense)
      f = FDAL_MASS_DENSE;
    else if (which == 0 && c->mass_cta_ws)
      f = FDAL_MASS_PCG_ONE_CTA, its = c->mass_its_m;
    else
      f = FDAL_MASS_PCG_KERNELS, its = which == 0 ? c->mass_its_m : c->mass_its_p;
  }
  if (form) *form = f;
  if (iterations) *iterations = its;
  if (interval_lo) *interval_lo = mc.mode ? mc.lo : 0.0;
  if (interval_hi) *interval_hi = mc.mode ? mc.hi : 0.0;
  if (verified_residual) *verified_residual = mc.mode ? mc.verified_resid : 0.0;
  return FDAL_OK;
}
int fdal_set_halo(fdal_ctx *c, int matrix_id, int level, int which, int64_t n_owned_cols, int64_t n_halo,
                  const int32_t *send_counts, const int32_t *send_idx, const int32_t *recv_counts) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (!send_counts || !recv_counts || n_owned_cols < 0 || n_halo < 0) return FDAL_ERR_INVALID;
  HostCsr *h = nullptr;
  if (matrix_id >= 0 && matrix_id < FDAL_MAT_COUNT) {
    h = &c->hmat[matrix_id];
  } else if (matrix_id >= FDAL_MAT_AMG_A && matrix_id <= FDAL_MAT_AMG_R) {
    if (which < 0 || which > 1 || level < 0 || level >= (int)c->amg[which].lev.size()) return FDAL_ERR_INVALID;
    AmgLevel &L = c->amg[which].lev[level];
    h = matrix_id == FDAL_MAT_AMG_A ? &L.hA : matrix_id == FDAL_MAT_AMG_P ? &L.hP : &L.hR;
  }
  if (!h || !h->set) {
    set_err(c, "fdal_set_halo: matrix %d (level %d) was not set", matrix_id, level);
    return FDAL_ERR_STATE;
  }
  if (n_owned_cols + n_halo != h->nc) {
    set_err(c, "fdal_set_halo: %lld owned + %lld halo columns != %lld columns", (long long)n_owned_cols,
            (long long)n_halo, (long long)h->nc);
    return FDAL_ERR_SHAPE;
  }
  int64_t ns = 0, nr = 0;
  for (int q = 0; q < c->nranks; ++q) {
    ns += send_counts[q];
    nr += recv_counts[q];
  }
  if (nr != n_halo) {
    set_err(c, "fdal_set_halo: recv_counts sum to %lld, expected %lld", (long long)nr, (long long)n_halo);
    return FDAL_ERR_SHAPE;
  }
  for (int64_t i = 0; i < ns; ++i)
    if (send_idx[i] < 0 || send_idx[i] >= n_owned_cols) {
      set_err(c, "fdal_set_halo: send index out of range");
      return FDAL_ERR_SHAPE;
    }
  h->has_plan = true;
  h->n_owned = n_owned_cols;
  h->n_halo = n_halo;
  h->send_counts.assign(send_counts, send_counts + c->nranks);
  h->recv_counts.assign(recv_counts, recv_counts + c->nranks);
  h->send_idx.assign(send_idx, send_idx + ns);
  return FDAL_OK;
}
int fdal_amg_set_coarse_range(fdal_ctx *c, int which, int64_t lo, int64_t hi) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (which < 0 || which > 1 || lo < 0 || hi < lo) return FDAL_ERR_INVALID;
  c->amg[which].c_lo = lo;
  c->amg[which].c_hi = hi;
  return FDAL_OK;
}
int fdal_amg_set_replicated_from(fdal_ctx *c, int which, int level) {
  CHECK_CTX(c);
  NOT_FINAL(c);
  if (which < 0 || which > 1 || level < 1) return FDAL_ERR_INVALID;
  c->amg[which].rep_from = level;
  return FDAL_OK;
}

}  // extern "C"
