/*
 * fdal.h — C ABI of the B200-native augmented-Lagrangian (AL) solve path.
 *
 * This is the drop-in boundary for the ONE hot path of
 * fdrmrc/fictitious_domain_AL_preconditioners: everything executed inside
 *     solver_fgmres.solve(AA, x, b, P_AL)
 * of immersed_laplace / stokes_immersed_boundary / elliptic_interface.
 * The host application (deal.II) keeps meshing, FE assembly and the
 * NonMatching coupling; it hands over CSR blocks, W^-1 and the AMG hierarchy
 * through the setters below and then calls either the per-vmult entry points
 * (parity mode, one host<->device round trip per call) or fdal_solve (the
 * whole outer Krylov solve stays on the device).
 *
 * Every entry point cites the reference interface it replaces
 * (file:line relative to the reference repository root).
 *
 * Conventions
 *   - plain pointers and sizes only, no C++ / torch types; never throws.
 *   - all functions return an int status (FDAL_OK == 0); a human readable
 *     message for the last failure is available from fdal_last_error().
 *   - matrices are CSR: int64 row_ptr[n_rows+1], int32 col[nnz], double
 *     val[nnz]; entries within a row may come in any order (deal.II stores the
 *     diagonal first, reference SURVEY A.7) — the library keeps the order it
 *     is given, so floating-point summation order is the caller's order.
 *     Exception: with fdal_config.block_size > 1 the matrix A and the finest
 *     AMG operator are stored as BSR with the block columns of a block row
 *     sorted ascending, so their row sums run in block-column order.
 *   - host-pointer calls copy their arguments; the caller's arrays only need
 *     to live for the duration of the call.
 *   - block vectors are passed as ONE contiguous array [block0|block1|block2]
 *     (deal.II BlockVector blocks are separate contiguous arrays; the adapter
 *     in INTEGRATION.md copies block-wise).
 *   - one host thread per context; a context owns its CUDA stream.
 */
#ifndef FDAL_H
#define FDAL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdal_ctx fdal_ctx;

/* ---- status codes ------------------------------------------------------ */
enum {
  FDAL_OK = 0,
  FDAL_ERR_INVALID = 1,        /* bad argument / enum value                  */
  FDAL_ERR_SHAPE = 2,          /* inconsistent matrix / vector sizes         */
  FDAL_ERR_ALLOC = 3,          /* host or device allocation failed           */
  FDAL_ERR_CUDA = 4,           /* CUDA runtime error                         */
  FDAL_ERR_STATE = 5,          /* call order (e.g. solve before finalize)    */
  FDAL_ERR_INNER_NO_CONVERGENCE = 6, /* inner CG hit max steps: deal.II would
                                  throw SolverControl::NoConvergence
                                  (immersed_laplace.cc:907-912)              */
  FDAL_ERR_OUTER_NO_CONVERGENCE = 7, /* outer FGMRES/MinRes hit max steps    */
  FDAL_ERR_MASS_NO_CONVERGENCE = 8,  /* Mp^-1 CG (stokes_immersed_boundary.cc
                                  :934-957) hit max steps                    */
  FDAL_ERR_NCCL = 9,
  FDAL_ERR_UNSUPPORTED = 10
};

/* ---- which saddle-point system / preconditioner ------------------------ */
enum {
  /* [[Ag,Ct],[C,0]], P = BlockPreconditionerAugmentedLagrangian
     (augmented_lagrangian_preconditioner.h:14-42, immersed_laplace.cc:891-944).
     blocks: (n | m) */
  FDAL_KIND_LAPLACE = 0,
  /* [[Ag,Bt,Ct],[B,0,0],[C,0,0]], P = ...AugmentedLagrangianStokes + FGMRES
     (augmented_lagrangian_preconditioner.h:44-79,
      stokes_immersed_boundary.cc:1000-1074). blocks: (n_u | n_p | m) */
  FDAL_KIND_STOKES = 1,
  /* same system, SPD block-diagonal P + MinRes
     (augmented_lagrangian_preconditioner.h:81-110,
      stokes_immersed_boundary.cc:1056-1064) */
  FDAL_KIND_STOKES_DIAG_MINRES = 2,
  /* elliptic interface 3x3, "ideal" AL: block CG on the 2x2 augmented block
     (augmented_lagrangian_preconditioner.h:115-164, elliptic_interface.cc:908-948).
     blocks: (n | m2 | m), m2 == m */
  FDAL_KIND_ELLIPTIC_IDEAL = 3,
  /* elliptic interface 3x3, modified AL: two scalar inner solves
     (augmented_lagrangian_preconditioner.h:168-238, elliptic_interface.cc:871-906) */
  FDAL_KIND_ELLIPTIC_MODIFIED = 4
};

/* ---- matrices the host exports ------------------------------------------ */
enum {
  FDAL_MAT_A = 0,   /* (1,1) block: stiffness / velocity block (n x n)        */
  FDAL_MAT_A2 = 1,  /* elliptic: immersed stiffness (beta2-beta1)(grad,grad)  */
  FDAL_MAT_BT = 2,  /* Stokes: B^T (n_u x n_p) — stokes_matrix.block(0,1)     */
  FDAL_MAT_B = 3,   /* Stokes: B (n_p x n_u)  — optional, else transposed Bt  */
  FDAL_MAT_CT = 4,  /* coupling_matrix as the reference stores it (n x m)     */
  FDAL_MAT_C = 5,   /* optional explicit C (m x n), else transposed Ct        */
  FDAL_MAT_M = 6,   /* immersed mass matrix (m x m)                           */
  FDAL_MAT_MP = 7,  /* Stokes pressure mass matrix (n_p x n_p)                */
  FDAL_MAT_COUNT = 8,
  /* only for fdal_set_halo: the matrices of one AMG level (which, level) */
  FDAL_MAT_AMG_A = 100,
  FDAL_MAT_AMG_P = 101,
  FDAL_MAT_AMG_R = 102
};

/* ---- W^-1 (augmented_lagrangian_preconditioner.h `invW`) --------------- */
enum {
  FDAL_WINV_DIAG = 0,            /* DiagonalMatrix given by fdal_set_diag     */
  FDAL_WINV_EXACT_M = 1,         /* M^-1      (UMFPACK in the reference)      */
  FDAL_WINV_EXACT_M_SQUARED = 2  /* M^-1 M^-1 (immersed_laplace.cc:875-876)   */
};
enum {
  FDAL_MPINV_CG_LUMPED = 0, /* CG(Mp; diag(Mp*1)^-1) stokes_immersed_boundary.cc:946-957 */
  FDAL_MPINV_EXACT = 1      /* UMFPACK in the reference (:960-962)             */
};
enum { FDAL_DIAG_W_INV = 0, FDAL_DIAG_MP_LUMPED_INV = 1 };
enum { FDAL_PREC_AMG = 0, FDAL_PREC_IDENTITY = 1 };
enum { FDAL_AMG_A11 = 0, FDAL_AMG_A22 = 1 };

/* ---- deal.II SolverControl family -------------------- */
enum {
  FDAL_CONTROL_SOLVER = 0,    /* success: value <= tol; failure: step >= max  */
  FDAL_CONTROL_REDUCTION = 1, /* also success: value < reduce * initial       */
  FDAL_CONTROL_ITERATION_NUMBER = 2 /* success when step >= max               */
};
typedef struct fdal_control {
  int32_t type;
  int32_t max_steps;
  double tol;
  double reduce;
} fdal_control;

typedef struct fdal_config {
  int32_t kind;        /* FDAL_KIND_*                                         */
  int32_t restart;     /* FGMRES max_basis_size: 30 default, 50 elliptic
                          (elliptic_interface.cc:863)                        */
  double gamma;        /* AL gamma (gamma_1); any 1/h, 1/h^2 scaling already
                          applied by the caller (immersed_laplace.cc:655-657,
                          elliptic_interface.cc:743-753)                     */
  double gamma2;       /* elliptic gamma_2                                    */
  double gamma_grad_div; /* Stokes gamma_grad_div                             */
  int32_t winv_mode;   /* FDAL_WINV_*                                         */
  int32_t mp_inv_mode; /* FDAL_MPINV_*                                        */
  int32_t aug_explicit; /* 1: operator form, A already contains the AL term
                           (immersed_laplace.cc:882)                         */
  int32_t grad_div_in_operator; /* 1: Aug += gamma_gd Bt Mp^-1 B
                           (stokes_immersed_boundary.cc:995)                 */
  int32_t inner_prec;  /* FDAL_PREC_AMG | FDAL_PREC_IDENTITY                  */
  int32_t device;      /* CUDA device ordinal                                 */
  int32_t use_graphs;  /* replay the V-cycle / CG iteration as CUDA graphs    */
  int32_t exact_mass_max_its; /* cap for the device mass solve (0 = default)  */
  int32_t block_size;  /* > 1: block0 unknowns are node-interleaved with this many
                          components per node (Stokes velocity dim, elasticity 3): A and
                          the finest AMG operator are stored as BSR (dim-blocked SpMV) */
  int32_t reserved0;
  fdal_control outer;  /* outer FGMRES / MinRes control                       */
  fdal_control inner;  /* inner CG on A_gamma (and on A22_gamma)              */
  fdal_control mass;   /* Mp^-1 CG control: (100, 1e-6)                       */
} fdal_config;

#define FDAL_MAX_HISTORY 1024
typedef struct fdal_solve_info {
  int32_t status;            /* FDAL_OK or the failure code                   */
  int32_t outer_iterations;  /* SolverControl::last_step()                    */
  int32_t inner_iterations;  /* sum over all inner A_gamma (A11) CG solves    */
  int32_t inner_iterations_a22; /* sum over A22_gamma solves (elliptic)       */
  int32_t inner_solves;      /* number of Aug_inv applications                */
  int32_t mass_iterations;   /* sum over Mp^-1 CG solves                      */
  int32_t n_history;         /* valid entries of residual_history             */
  int32_t reserved;          /* CUDA graph launches issued by the solve              */
  double initial_residual;
  double final_residual;
  double solve_ms;           /* device time of the solve (CUDA events)        */
  int64_t kernel_launches;   /* kernels of this library launched in the solve */
  double residual_history[FDAL_MAX_HISTORY];
} fdal_solve_info;

/* ---- lifetime ------------------------------------------------------------ */
int fdal_create(fdal_ctx **out, const fdal_config *cfg);
void fdal_destroy(fdal_ctx *ctx);
const char *fdal_last_error(const fdal_ctx *ctx);
const char *fdal_version(void);

/* ---- setup: replaces the linear_operator(...) wrappers ------------------- *
 * immersed_laplace.cc:638-643, stokes_immersed_boundary.cc:923-929,
 * elliptic_interface.cc:680-687 */
int fdal_set_csr(fdal_ctx *ctx, int matrix_id, int64_t n_rows, int64_t n_cols,
                 int64_t nnz, const int64_t *row_ptr, const int32_t *col,
                 const double *val);
/* DiagonalMatrix payloads: inverse_squares / inv_diagonal
 * (immersed_laplace.cc:853-873, stokes_immersed_boundary.cc:946-983,
 *  elliptic_interface.cc:705-728, utilities.h:348-374) */
int fdal_set_diag(fdal_ctx *ctx, int diag_id, int64_t n, const double *d);

/* AMG hierarchy built by TrilinosWrappers::PreconditionAMG::initialize
 * (immersed_laplace.cc:704,833; utilities.h:308-317,729-733;
 *  elliptic_interface.cc:824-850).  Level 0 = finest.  `P` maps level+1 ->
 * level (n_l x n_{l+1}); `R` may be NULL (then R = P^T).  inv_diag may be NULL
 * (then 1/diag(A)).  Chebyshev: degree sweeps, eigenvalue interval
 * [lambda_max/eig_ratio, 1.1*lambda_max].
 * The last level set is the coarsest: it only needs A (direct solve). */
typedef struct fdal_csr_view {
  int64_t n_rows, n_cols, nnz;
  const int64_t *row_ptr;
  const int32_t *col;
  const double *val;
} fdal_csr_view;
int fdal_amg_set_level(fdal_ctx *ctx, int which, int level,
                       const fdal_csr_view *A, const fdal_csr_view *P,
                       const fdal_csr_view *R, const double *inv_diag,
                       double lambda_max, int cheb_degree, double eig_ratio);
int fdal_amg_set_coarse(fdal_ctx *ctx, int which, int level,
                        const fdal_csr_view *A);

/* upload, build transposes / kernel schedules, allocate the Krylov workspace */
int fdal_finalize(fdal_ctx *ctx);

/* sizes of the blocks of the system vector, n_blocks in {2,3} */
int fdal_block_sizes(const fdal_ctx *ctx, int64_t sizes[3], int *n_blocks);

/* ---- per-vmult entry points (host pointers) ------------------------------ */
/* SparseMatrix::vmult / Tvmult on an exported block (K1/K2/K6/K7) */
int fdal_spmv(fdal_ctx *ctx, int matrix_id, int transpose, const double *x,
              double *y);
/* Aug.vmult: y = A x + gamma Ct W^-1 C x  (immersed_laplace.cc:880-884,
 * stokes_immersed_boundary.cc:991-995, elliptic_interface.cc:807); which =
 * FDAL_AMG_A11 for A_gamma / A11_gamma, FDAL_AMG_A22 for A22_gamma (:810) */
int fdal_apply_aug(fdal_ctx *ctx, int which, const double *x, double *y);
/* AA.vmult / system_operator.vmult (immersed_laplace.cc:891-892,
 * stokes_immersed_boundary.cc:1000-1003, elliptic_interface.cc:816-819) */
int fdal_apply_system(fdal_ctx *ctx, const double *x, double *y);
/* invW.vmult */
int fdal_apply_winv(fdal_ctx *ctx, const double *x, double *y);
/* Mp_inv.vmult (stokes_immersed_boundary.cc:931-963) */
int fdal_apply_mp_inv(fdal_ctx *ctx, const double *x, double *y, int *its);
/* TrilinosWrappers::PreconditionAMG::vmult — one V-cycle */
int fdal_apply_amg(fdal_ctx *ctx, int which, const double *r, double *z);
/* Aug_inv.vmult = inverse_operator(Aug, SolverCG, AMG) from a zero guess
 * (immersed_laplace.cc:911-912, stokes_immersed_boundary.cc:1045,
 *  elliptic_interface.cc:895-898) */
int fdal_apply_aug_inv(fdal_ctx *ctx, int which, const double *b, double *x,
                       int *its);
/* P.vmult of the configured AL preconditioner
 * (augmented_lagrangian_preconditioner.h:28-34, 62-70, 95-103, 130-156, 186-229).
 * inner_its[0] = A11 CG iterations, inner_its[1] = A22 CG iterations. */
int fdal_apply_prec(fdal_ctx *ctx, const double *u, double *v, int inner_its[2]);
/* rhs0 = f + gamma Ct W^-1 g (immersed_laplace.cc:899-905,
 * stokes_immersed_boundary.cc:1011-1018): in-place on the block rhs */
int fdal_augment_rhs(fdal_ctx *ctx, double *rhs_inout);

/* ---- the solve: replaces solver_fgmres.solve(AA, x, b, P) ----------------
 * immersed_laplace.cc:943-944, stokes_immersed_boundary.cc:1063-1074,
 * elliptic_interface.cc:905,947.  x_inout holds the initial guess. */
int fdal_solve(fdal_ctx *ctx, const double *rhs, double *x_inout,
               fdal_solve_info *info);

/* ---- device-pointer variants (inputs already resident in HBM) ------------ */
int fdal_solve_dev(fdal_ctx *ctx, const double *d_rhs, double *d_x_inout,
                   fdal_solve_info *info);
int fdal_apply_aug_dev(fdal_ctx *ctx, int which, const double *d_x, double *d_y);
int fdal_apply_amg_dev(fdal_ctx *ctx, int which, const double *d_r, double *d_z);
int fdal_spmv_dev(fdal_ctx *ctx, int matrix_id, int transpose, const double *d_x,
                  double *d_y);

/* ---- measurement: CUDA-event timing of one kernel family on the ctx stream */
enum {
  FDAL_TIME_SPMV_A = 0,     /* y = A x                                        */
  FDAL_TIME_AUG = 1,        /* y = A x + gamma Ct W^-1 C x (diag W^-1)        */
  FDAL_TIME_VCYCLE = 2,     /* one AMG V-cycle (which = A11)                  */
  FDAL_TIME_CHEB_FINE = 3,  /* one fused Chebyshev step on the finest level   */
  FDAL_TIME_DOT = 4,        /* fused dot on an N-vector                       */
  FDAL_TIME_MULTIDOT = 5,   /* V^T w with `param` basis vectors               */
  FDAL_TIME_AXPY = 6
};
/* runs `reps` launches after `warmup`, optionally flushing L2 between them;
 * returns average ms and the algorithmic bytes of one launch */
int fdal_time_kernel(fdal_ctx *ctx, int what, int param, int warmup, int reps,
                     int flush_l2, double *avg_ms, double *algorithmic_bytes,
                     int64_t *launches_per_rep);

/* ---- multi-GPU (one process per GPU) -------------------------------------
 * No reference counterpart (the reference is serial).  Rows of the
 * background (and pressure) unknowns are partitioned into contiguous ranges; the
 * multiplier block and the immersed blocks are replicated.  Call order:
 * fdal_create, fdal_comm_init, fdal_set_csr / fdal_amg_set_level with the LOCAL rows
 * (columns numbered [owned | halo]), fdal_set_halo for every matrix that has a halo,
 * fdal_amg_set_coarse with the full (replicated) coarsest operator +
 * fdal_amg_set_coarse_range, fdal_finalize.  Vectors passed to the apply / solve
 * entry points are then the rank-local [block0_loc | block1_loc | replicated tail]. */
int fdal_nccl_unique_id(char id_out[128]);
int fdal_comm_init(fdal_ctx *ctx, const char id[128], int rank, int n_ranks);
/* halo plan of a row-partitioned matrix whose columns are numbered
 * [owned | halo]: for each peer the owned entries to send (send_idx, grouped by
 * destination rank) and the number of halo entries received (halo entries are
 * ordered by owner rank).  matrix_id: FDAL_MAT_* or FDAL_MAT_AMG_{A,P,R} with
 * (which, level). */
int fdal_set_halo(fdal_ctx *ctx, int matrix_id, int level, int which,
                  int64_t n_owned_cols, int64_t n_halo, const int32_t *send_counts,
                  const int32_t *send_idx, const int32_t *recv_counts);
/* Agglomeration: levels >= `level` of hierarchy `which` are replicated on every rank (full
 * matrices, no halo plans; default: only the coarsest operator).  The last partitioned level's P
 * then has GLOBAL columns of level `level` and its R has this rank's rows [lo, hi) of that level
 * (fdal_amg_set_coarse_range); the restricted residual is all-gathered, everything below runs
 * redundantly without communication. */
int fdal_amg_set_replicated_from(fdal_ctx *ctx, int which, int level);
/* rows [lo, hi) of the first replicated level belong to this rank */
int fdal_amg_set_coarse_range(fdal_ctx *ctx, int which, int64_t lo, int64_t hi);
/* 0: single rank; 1: NCCL send/recv + all-reduce per exchange (fallback, FDAL_COMM=nccl);
 * 2: peer channels — every rank maps the others' exchange arena (cudaIpc) and halos, scalar and
 * vector all-reduces and the coarse all-gather are NVLink stores + flags issued by this
 * library's own kernels (fused into the reductions' last block / the SpMV's boundary chunks) */
int fdal_comm_mode(const fdal_ctx *ctx);

/* How the exact mass inverses (the device replacement of SparseDirectUMFPACK::vmult: immersed_laplace.cc:864,
 * 875-876; stokes_immersed_boundary.cc:960-962, 981-985; elliptic_interface.cc:719-720, 736-737) are applied,
 * decided at fdal_finalize.  which = 0: the multiplier mass matrix M (W^-1 = M^-1 or M^-2), 1: the pressure
 * mass matrix Mp.  *iterations: the fixed iteration count of one solve (0 for the dense form);
 * [*interval_lo, *interval_hi]: the spectral interval of D^-1 M the Chebyshev coefficients were built on;
 * *verified_residual: |b - M x| / |b| of the Chebyshev solve on the calibration right-hand side (it must be
 * <= 2e-14 for the Chebyshev form to be selected).  Any out pointer may be NULL. */
enum {
  FDAL_MASS_NONE = 0,            /* no exact mass solve in this configuration */
  FDAL_MASS_PCG_KERNELS = 1,     /* fixed-count Jacobi-PCG, four kernels per iteration, one CUDA graph */
  FDAL_MASS_PCG_ONE_CTA = 2,     /* the same iteration inside one CTA (m <= 16384) */
  FDAL_MASS_DENSE = 3,           /* dense W^-1 GEMV (m <= 4096) */
  FDAL_MASS_CHEB_KERNELS = 4,    /* fixed-count Chebyshev, one fused SpMV kernel per iteration */
  FDAL_MASS_CHEB_PERSISTENT = 5  /* fixed-count Chebyshev in one persistent kernel, matrix staged in shared memory */
};
int fdal_mass_solver_info(const fdal_ctx *ctx, int which, int32_t *form, int32_t *iterations, double *interval_lo,
                          double *interval_hi, double *verified_residual);

/* ---- setup phase on the device: CSR -> BSR conversion (part of SURVEY 8(f) N2: the data preparation the
 * reference leaves to deal.II / Trilinos on the host) ---------------------------------------------------
 * With fdal_config.block_size = 2 or 3 fdal_finalize stores the velocity / elasticity block and the finest AMG
 * operator as BSR.  The conversion runs on the device on the scalar CSR arrays that were uploaded
 * anyway (csrc/bsr_build.cu: per-warp shared-memory hash set of the block columns, prefix sum, sorted fill);
 * FDAL_HOST_BSR=1 keeps the OpenMP conversion on the host.  fdal_bsr_conversions reports where the conversions of
 * a finalized context ran.  fdal_csr_to_bsr is the same device routine with host arrays in and out (tests, tools):
 * returns the number of blocks (>= 0), -1000 when the matrix is not blocked (explicit zeros > max_fill x nnz, or
 * sizes beyond the 32-bit block pointers), or -FDAL_ERR_* .  brow_ptr_out has n_rows / block_size + 1 entries and is
 * always written; bcol_out / bval_out (block_capacity blocks) are written when they are large enough, so a first
 * call with block_capacity = 0 sizes the second.  Blocks are contiguous, row-major, block columns ascending. */
int fdal_bsr_conversions(const fdal_ctx *ctx, int32_t *on_device, int32_t *on_host);
int64_t fdal_csr_to_bsr(int device, int64_t n_rows, int64_t nnz, const int64_t *row_ptr, const int32_t *col,
                        const double *val, int32_t block_size, double max_fill, int32_t *brow_ptr_out,
                        int64_t block_capacity, int32_t *bcol_out, double *bval_out);

/* ---- setup phase on the device -----------------------------------
 * Operator-form AL term (immersed_laplace.cc:659-702 with the particles of utilities.h:755-837;
 * nitsche_bcs.cc:517-572):  A += sum_q weight[q] * phi_q phi_q^T  scattered into the CSR values of the
 * stiffness matrix on the device.  The host (deal.II Particles / FEValues) supplies, per immersed
 * quadrature point q, the dofs_per_cell dof indices of the background cell it lies in (point_dofs, a
 * negative index skips that local dof) and the shape-function values there (point_phi);
 * weight[q] = gamma * JxW_q.  val_inout is updated in place (row_ptr / col unchanged; rows need not be
 * sorted).  Returns FDAL_ERR_SHAPE, with the count in *n_missing_out, if the sparsity pattern lacks an
 * entry a point needs.  Stand-alone: no fdal_ctx involved. */
int fdal_assemble_al_term(int device, int64_t n_rows, const int64_t *row_ptr, const int32_t *col,
                          double *val_inout, int64_t n_points, int32_t dofs_per_cell,
                          const int32_t *point_dofs, const double *point_phi, const double *weight,
                          int64_t *n_missing_out);

#ifdef __cplusplus
}
#endif
#endif /* FDAL_H */
