/*
 * fdal_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A scalar C restatement of the reference's augmented-Lagrangian solve path
 * (fdrmrc/fictitious_domain_AL_preconditioners).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libfdal.so, CUDA) never does.
 *
 * PARITY: PARTLY PINNED.
 *   PINNED against the reference's own code: the algebra of the five AL
 *   preconditioner vmults (augmented_lagrangian_preconditioner.h:28-34, 62-70,
 *   95-103, 130-156, 225-228).  That header is compiled UNMODIFIED from
 *   /root/reference against stand-in deal.II vector/LinearOperator types
 *   (oracle/ref_harness, `make ref` -> oracle/_ref/libref_prec.so) and its vmults
 *   are run on seeded inputs; fdalo_apply_prec is bit-identical to them
 *   (tests/test_reference_pinning.py) and the outputs are committed as
 *   tests/golden/ref_prec_vectors.npz.
 *   UNPINNED (restated from published algorithms, checked only by algebraic
 *   known answers, tests/test_oracle_known_answers.py): everything whose
 *   arithmetic lives in dependencies that are neither vendored under the
 *   reference tree nor installable here — deal.II >= 9.6 (SolverControl /
 *   ReductionControl / IterationNumberControl rules, SolverCG, SolverFGMRES,
 *   SolverMinRes, inverse_operator, SparseMatrix::vmult/Tvmult; SURVEY.md App.
 *   A.1-A.5, A.7), Trilinos ML >= 14.4 (V-cycle with Chebyshev smoothing, App.
 *   A.6) and UMFPACK — and the operator definitions inside the three
 *   applications (immersed_laplace.cc:880-905, stokes_immersed_boundary.cc:
 *   931-1018, elliptic_interface.cc:693-819), which need all of deal.II to build.
 *   The reference ships no tests, golden vectors or committed AL iteration counts.
 *
 * The API mirrors include/fdal.h with the prefix fdalo_ so the same Python
 * binding drives both sides of a parity test.  Exact mass inverses are direct
 * solves with SuperLU factors handed in by the test harness (the stand-in for
 * SparseDirectUMFPACK::initialize); the AMG coarse solve is a dense LU.
 *
 * Build: make -C oracle      (gcc -O3 -fopenmp; thread count via
 * fdalo_set_num_threads, default 1 = how the reference ships:
 * MPI_InitFinalize(argc, argv, 1), immersed_laplace.cc:1048).
 */
#include "../include/fdal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OMP_MIN 20000

typedef struct {
  int64_t nr, nc, nnz;
  int64_t *rp;
  int32_t *ci;
  double *v;
  int set;
} csr;

typedef struct {
  int64_t n;
  int64_t *Lp, *Up;
  int32_t *Li, *Ui;
  double *Lx, *Ux;
  int32_t *perm_r, *perm_c;
  double *w;
  int set;
} lu_t;

typedef struct {
  csr A, P, R;
  double *inv_diag;
  double lmax, ratio;
  int degree;
  double *x, *b, *r, *d;
} amg_level;

#define MAX_LEVELS 24
typedef struct {
  int nlev; /* including coarsest */
  amg_level lev[MAX_LEVELS];
  int64_t cn;
  double *clu;
  int *cpiv;
  int ready;
} amg_t;

typedef struct {
  fdal_control c;
  double initial, reduced_tol;
  int last_step;
  double last_value;
} control_state;

struct fdalo_ctx {
  fdal_config cfg;
  csr mat[FDAL_MAT_COUNT];
  double *winv_diag;
  int64_t winv_n;
  double *mp_lumped_inv;
  int64_t mp_n;
  lu_t lu_m, lu_mp;
  amg_t amg[2];
  int finalized;
  int64_t n0, n1, n2, N;
  int nblocks;
  /* scratch */
  double *tm0, *tm1, *tm2, *tn0, *tn1, *tp0, *tp1, *tb0;
  /* counters */
  int its_a11, its_a22, its_mass, n_inner_solves;
  int fail;
  char err[512];
};
typedef struct fdalo_ctx fdalo_ctx;

static int g_threads = 1;
void fdalo_set_num_threads(int n) {
  g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(g_threads);
#endif
}
int fdalo_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ vectors */
static double *dalloc(int64_t n) { return (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double)); }

static double vdot(int64_t n, const double *a, const double *b) {
  double s = 0;
#pragma omp parallel for reduction(+ : s) if (n > OMP_MIN) schedule(static)
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static void vcopy(int64_t n, const double *a, double *y) { memcpy(y, a, (size_t)n * sizeof(double)); }
static void vzero(int64_t n, double *y) { memset(y, 0, (size_t)n * sizeof(double)); }
/* y += a x */
static void vaxpy(int64_t n, double a, const double *x, double *y) {
#pragma omp parallel for if (n > OMP_MIN) schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}
/* y = s y + a x  (deal.II Vector::sadd) */
static void vsadd(int64_t n, double s, double a, const double *x, double *y) {
#pragma omp parallel for if (n > OMP_MIN) schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = s * y[i] + a * x[i];
}
static void vscale(int64_t n, double a, double *y) {
#pragma omp parallel for if (n > OMP_MIN) schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] *= a;
}

/* ------------------------------------------------------------------ CSR */
static void csr_free(csr *m) {
  free(m->rp);
  free(m->ci);
  free(m->v);
  memset(m, 0, sizeof(*m));
}
static int csr_copy(csr *m, int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp, const int32_t *ci,
                    const double *v) {
  csr_free(m);
  m->nr = nr;
  m->nc = nc;
  m->nnz = nnz;
  m->rp = (int64_t *)malloc((size_t)(nr + 1) * sizeof(int64_t));
  m->ci = (int32_t *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t));
  m->v = (double *)malloc((size_t)(nnz > 0 ? nnz : 1) * sizeof(double));
  if (!m->rp || !m->ci || !m->v) return FDAL_ERR_ALLOC;
  memcpy(m->rp, rp, (size_t)(nr + 1) * sizeof(int64_t));
  if (nnz) {
    memcpy(m->ci, ci, (size_t)nnz * sizeof(int32_t));
    memcpy(m->v, v, (size_t)nnz * sizeof(double));
  }
  m->set = 1;
  return FDAL_OK;
}
static int csr_from_view(csr *m, const fdal_csr_view *v) {
  return csr_copy(m, v->n_rows, v->n_cols, v->nnz, v->row_ptr, v->col, v->val);
}
/* y = beta*y + alpha*A x  (SparseMatrix::vmult: row gather) */
static void spmv(const csr *A, const double *x, double *y, double alpha, double beta) {
  const int64_t nr = A->nr;
#pragma omp parallel for if (nr > OMP_MIN / 4) schedule(static)
  for (int64_t i = 0; i < nr; ++i) {
    double s = 0;
    for (int64_t k = A->rp[i]; k < A->rp[i + 1]; ++k) s += A->v[k] * x[A->ci[k]];
    y[i] = (beta == 0.0 ? 0.0 : beta * y[i]) + alpha * s;
  }
}
/* y = beta*y + alpha*A^T x  (SparseMatrix::Tvmult: row scatter, serial like deal.II) */
static void spmv_t(const csr *A, const double *x, double *y, double alpha, double beta) {
  if (beta == 0.0)
    vzero(A->nc, y);
  else if (beta != 1.0)
    vscale(A->nc, beta, y);
  for (int64_t i = 0; i < A->nr; ++i) {
    const double xi = alpha * x[i];
    if (xi == 0.0) continue;
    for (int64_t k = A->rp[i]; k < A->rp[i + 1]; ++k) y[A->ci[k]] += A->v[k] * xi;
  }
}

/* ------------------------------------------------------------------ LU (SuperLU factors) */
static void lu_free(lu_t *f) {
  free(f->Lp);
  free(f->Up);
  free(f->Li);
  free(f->Ui);
  free(f->Lx);
  free(f->Ux);
  free(f->perm_r);
  free(f->perm_c);
  free(f->w);
  memset(f, 0, sizeof(*f));
}
/* scipy.sparse.linalg.splu: Pr A Pc = L U with (Pr b)[perm_r[i]] = b[i] and
 * x[i] = y[perm_c[i]]; checked against lu.solve in tests/test_oracle_known_answers.py */
static void lu_solve(const lu_t *f, const double *b, double *x) {
  const int64_t n = f->n;
  double *w = f->w;
  for (int64_t i = 0; i < n; ++i) w[f->perm_r[i]] = b[i];
  /* L (CSC, unit diagonal stored) forward */
  for (int64_t j = 0; j < n; ++j) {
    double dj = 1.0;
    for (int64_t k = f->Lp[j]; k < f->Lp[j + 1]; ++k)
      if (f->Li[k] == j) dj = f->Lx[k];
    w[j] /= dj;
    const double wj = w[j];
    for (int64_t k = f->Lp[j]; k < f->Lp[j + 1]; ++k)
      if (f->Li[k] > j) w[f->Li[k]] -= f->Lx[k] * wj;
  }
  /* U (CSC) backward */
  for (int64_t j = n - 1; j >= 0; --j) {
    double dj = 1.0;
    for (int64_t k = f->Up[j]; k < f->Up[j + 1]; ++k)
      if (f->Ui[k] == j) dj = f->Ux[k];
    w[j] /= dj;
    const double wj = w[j];
    for (int64_t k = f->Up[j]; k < f->Up[j + 1]; ++k)
      if (f->Ui[k] < j) w[f->Ui[k]] -= f->Ux[k] * wj;
  }
  for (int64_t i = 0; i < n; ++i) x[i] = w[f->perm_c[i]];
}

/* ------------------------------------------------------------------ controls */
enum { ST_ITERATE = 0, ST_SUCCESS = 1, ST_FAILURE = 2 };
static int control_check(control_state *s, int step, double val) {
  s->last_step = step;
  s->last_value = val;
  if (s->c.type == FDAL_CONTROL_REDUCTION) {
    if (step == 0) {
      s->initial = val;
      s->reduced_tol = val * s->c.reduce;
    }
    if (val < s->reduced_tol) return ST_SUCCESS;
  } else if (s->c.type == FDAL_CONTROL_ITERATION_NUMBER) {
    if (step >= s->c.max_steps) return ST_SUCCESS;
  }
  if (val <= s->c.tol) return ST_SUCCESS;
  if (step >= s->c.max_steps || isnan(val)) return ST_FAILURE;
  return ST_ITERATE;
}

/* ------------------------------------------------------------------ AMG V-cycle */
static void cheb(const amg_level *L, const double *b, double *x, int zero_guess) {
  const int64_t n = L->A.nr;
  const double beta = 1.1 * L->lmax, alpha = L->lmax / L->ratio;
  const double delta = 0.5 * (beta - alpha), theta = 0.5 * (beta + alpha);
  const double s1 = theta / delta;
  double rho = 1.0 / s1;
  double *d = L->d, *r = L->r;
  const double *id = L->inv_diag;
  if (zero_guess) {
    for (int64_t i = 0; i < n; ++i) {
      d[i] = id[i] * b[i] / theta;
      x[i] = d[i];
    }
  } else {
    vcopy(n, b, r);
    spmv(&L->A, x, r, -1.0, 1.0);
    for (int64_t i = 0; i < n; ++i) {
      d[i] = id[i] * r[i] / theta;
      x[i] += d[i];
    }
  }
  for (int k = 1; k < L->degree; ++k) {
    const double rho1 = 1.0 / (2.0 * s1 - rho);
    vcopy(n, b, r);
    spmv(&L->A, x, r, -1.0, 1.0);
    const double c1 = rho1 * rho, c2 = 2.0 * rho1 / delta;
#pragma omp parallel for if (n > OMP_MIN) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      d[i] = c1 * d[i] + c2 * id[i] * r[i];
      x[i] += d[i];
    }
    rho = rho1;
  }
}
static void dense_lu_solve(int64_t n, const double *lu, const int *piv, const double *b, double *x) {
  for (int64_t i = 0; i < n; ++i) x[i] = b[i];
  /* whole rows (multipliers included) were exchanged during factorisation,
   * getrf style: apply all row swaps first, then substitute */
  for (int64_t i = 0; i < n; ++i) {
    const int p = piv[i];
    if (p != i) {
      double t = x[i];
      x[i] = x[p];
      x[p] = t;
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    const double xi = x[i];
    if (xi != 0.0)
      for (int64_t r = i + 1; r < n; ++r) x[r] -= lu[r * n + i] * xi;
  }
  for (int64_t i = n - 1; i >= 0; --i) {
    double s = x[i];
    for (int64_t c = i + 1; c < n; ++c) s -= lu[i * n + c] * x[c];
    x[i] = s / lu[i * n + i];
  }
}
static void vcycle(amg_t *g, int l, const double *b, double *x) {
  amg_level *L = &g->lev[l];
  if (l == g->nlev - 1) {
    dense_lu_solve(g->cn, g->clu, g->cpiv, b, x);
    return;
  }
  amg_level *C = &g->lev[l + 1];
  const int64_t n = L->A.nr;
  cheb(L, b, x, 1);
  vcopy(n, b, L->r);
  spmv(&L->A, x, L->r, -1.0, 1.0);
  if (L->R.set)
    spmv(&L->R, L->r, C->b, 1.0, 0.0);
  else
    spmv_t(&L->P, L->r, C->b, 1.0, 0.0);
  vcycle(g, l + 1, C->b, C->x);
  spmv(&L->P, C->x, x, 1.0, 1.0);
  cheb(L, b, x, 0);
}
static void amg_apply(amg_t *g, const double *r, double *z) { vcycle(g, 0, r, z); }

static int amg_prepare(amg_t *g, char *err) {
  if (g->nlev == 0) return FDAL_OK;
  for (int l = 0; l < g->nlev; ++l) {
    amg_level *L = &g->lev[l];
    if (!L->A.set) {
      snprintf(err, 512, "AMG level %d missing", l);
      return FDAL_ERR_STATE;
    }
    const int64_t n = L->A.nr;
    if (!L->inv_diag) {
      L->inv_diag = dalloc(n);
      for (int64_t i = 0; i < n; ++i) {
        double dsum = 0;
        for (int64_t k = L->A.rp[i]; k < L->A.rp[i + 1]; ++k)
          if (L->A.ci[k] == i) dsum += L->A.v[k];
        L->inv_diag[i] = 1.0 / dsum;
      }
    }
    L->x = dalloc(n);
    L->b = dalloc(n);
    L->r = dalloc(n);
    L->d = dalloc(n);
  }
  /* dense LU with partial pivoting of the coarsest operator (Amesos-KLU stand-in) */
  amg_level *C = &g->lev[g->nlev - 1];
  const int64_t n = C->A.nr;
  g->cn = n;
  g->clu = dalloc(n * n);
  g->cpiv = (int *)malloc((size_t)n * sizeof(int));
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = C->A.rp[i]; k < C->A.rp[i + 1]; ++k) g->clu[i * n + C->A.ci[k]] += C->A.v[k];
  for (int64_t k = 0; k < n; ++k) {
    int64_t p = k;
    double best = fabs(g->clu[k * n + k]);
    for (int64_t r = k + 1; r < n; ++r)
      if (fabs(g->clu[r * n + k]) > best) {
        best = fabs(g->clu[r * n + k]);
        p = r;
      }
    g->cpiv[k] = (int)p;
    if (best == 0.0) {
      snprintf(err, 512, "singular coarse AMG operator at column %ld", (long)k);
      return FDAL_ERR_INVALID;
    }
    if (p != k)
      for (int64_t c = 0; c < n; ++c) {
        double t = g->clu[k * n + c];
        g->clu[k * n + c] = g->clu[p * n + c];
        g->clu[p * n + c] = t;
      }
    const double piv = g->clu[k * n + k];
    for (int64_t r = k + 1; r < n; ++r) {
      const double f = g->clu[r * n + k] / piv;
      g->clu[r * n + k] = f;
      if (f != 0.0)
        for (int64_t c = k + 1; c < n; ++c) g->clu[r * n + c] -= f * g->clu[k * n + c];
    }
  }
  g->ready = 1;
  return FDAL_OK;
}

/* ------------------------------------------------------------------ generic CG */
typedef void (*apply_fn)(fdalo_ctx *, int, const double *, double *);

static int cg_solve(fdalo_ctx *c, int64_t n, apply_fn op, int op_arg, apply_fn prec, int prec_arg,
                    const fdal_control *ctl, const double *b, double *x, int *its_out) {
  /* inverse_operator re-zeros dst: x = 0 => r = b */
  double *r = dalloc(n), *z = dalloc(n), *p = dalloc(n), *v = dalloc(n);
  control_state cs;
  memset(&cs, 0, sizeof(cs));
  cs.c = *ctl;
  vzero(n, x);
  vcopy(n, b, r);
  double res = sqrt(vdot(n, r, r));
  int st = control_check(&cs, 0, res);
  int it = 0;
  double rho_old = 0;
  while (st == ST_ITERATE) {
    ++it;
    if (prec)
      prec(c, prec_arg, r, z);
    else
      vcopy(n, r, z);
    const double rho = vdot(n, r, z);
    if (it > 1) {
      const double beta = rho / rho_old;
      vsadd(n, beta, 1.0, z, p);
    } else
      vcopy(n, z, p);
    op(c, op_arg, p, v);
    const double alpha = rho / vdot(n, p, v);
    vaxpy(n, alpha, p, x);
    vaxpy(n, -alpha, v, r);
    res = sqrt(fabs(vdot(n, r, r)));
    rho_old = rho;
    st = control_check(&cs, it, res);
  }
  free(r);
  free(z);
  free(p);
  free(v);
  *its_out = it;
  return st == ST_SUCCESS ? FDAL_OK : FDAL_ERR_INNER_NO_CONVERGENCE;
}

/* ------------------------------------------------------------------ the operators of the path */
static const csr *M_(fdalo_ctx *c, int id) { return &c->mat[id]; }

/* C x: transpose_operator(linear_operator(coupling_matrix)) -> Tvmult on Ct
 * (immersed_laplace.cc:640-641); an explicit C is used when given */
static void apply_C(fdalo_ctx *c, const double *x, double *y) {
  if (c->mat[FDAL_MAT_C].set)
    spmv(M_(c, FDAL_MAT_C), x, y, 1.0, 0.0);
  else
    spmv_t(M_(c, FDAL_MAT_CT), x, y, 1.0, 0.0);
}
static void apply_B(fdalo_ctx *c, const double *x, double *y) {
  if (c->mat[FDAL_MAT_B].set)
    spmv(M_(c, FDAL_MAT_B), x, y, 1.0, 0.0);
  else
    spmv_t(M_(c, FDAL_MAT_BT), x, y, 1.0, 0.0);
}
/* invW (immersed_laplace.cc:849-878, stokes_immersed_boundary.cc:966-985,
 * elliptic_interface.cc:693-739) */
static void apply_winv(fdalo_ctx *c, const double *x, double *y) {
  const int64_t m = c->winv_n;
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    for (int64_t i = 0; i < m; ++i) y[i] = c->winv_diag[i] * x[i];
  } else if (c->cfg.winv_mode == FDAL_WINV_EXACT_M) {
    lu_solve(&c->lu_m, x, y);
  } else {
    double *t = dalloc(m);
    lu_solve(&c->lu_m, x, t);
    lu_solve(&c->lu_m, t, y);
    free(t);
  }
}
static void mp_op(fdalo_ctx *c, int a, const double *x, double *y) {
  (void)a;
  spmv(M_(c, FDAL_MAT_MP), x, y, 1.0, 0.0);
}
static void mp_prec(fdalo_ctx *c, int a, const double *x, double *y) {
  (void)a;
  for (int64_t i = 0; i < c->mp_n; ++i) y[i] = c->mp_lumped_inv[i] * x[i];
}
/* Mp_inv (stokes_immersed_boundary.cc:931-963) */
static void apply_mp_inv(fdalo_ctx *c, const double *x, double *y) {
  if (c->cfg.mp_inv_mode == FDAL_MPINV_EXACT) {
    lu_solve(&c->lu_mp, x, y);
  } else {
    int its = 0;
    int st = cg_solve(c, c->mp_n, mp_op, 0, mp_prec, 0, &c->cfg.mass, x, y, &its);
    c->its_mass += its;
    if (st != FDAL_OK && !c->fail) c->fail = FDAL_ERR_MASS_NO_CONVERGENCE;
  }
}

/* Aug / A11_aug / A22_aug (immersed_laplace.cc:880-884,
 * stokes_immersed_boundary.cc:991-995, elliptic_interface.cc:807-810) */
static void apply_aug(fdalo_ctx *c, int which, const double *x, double *y) {
  const int64_t m = c->winv_n;
  if (which == FDAL_AMG_A11) {
    spmv(M_(c, FDAL_MAT_A), x, y, 1.0, 0.0);
    if (!c->cfg.aug_explicit) {
      double *t = dalloc(m), *u = dalloc(m);
      apply_C(c, x, t);
      apply_winv(c, t, u);
      spmv(M_(c, FDAL_MAT_CT), u, y, c->cfg.gamma, 1.0);
      free(t);
      free(u);
    }
    if ((c->cfg.kind == FDAL_KIND_STOKES || c->cfg.kind == FDAL_KIND_STOKES_DIAG_MINRES) &&
        c->cfg.grad_div_in_operator) {
      double *t = dalloc(c->n1), *u = dalloc(c->n1);
      apply_B(c, x, t);
      apply_mp_inv(c, t, u);
      spmv(M_(c, FDAL_MAT_BT), u, y, c->cfg.gamma_grad_div, 1.0);
      free(t);
      free(u);
    }
  } else {
    /* A22_aug = A_omega2 + gamma_2 * M * invW * M */
    double *t = dalloc(m), *u = dalloc(m);
    spmv(M_(c, FDAL_MAT_A2), x, y, 1.0, 0.0);
    spmv(M_(c, FDAL_MAT_M), x, t, 1.0, 0.0);
    apply_winv(c, t, u);
    spmv(M_(c, FDAL_MAT_M), u, y, c->cfg.gamma2, 1.0);
    free(t);
    free(u);
  }
}
/* A12_aug = -gamma_1 Ct invW M ; A21_aug = -gamma_2 M invW C (elliptic_interface.cc:811-813) */
static void apply_a12_add(fdalo_ctx *c, const double *x2, double *y0) {
  const int64_t m = c->winv_n;
  double *t = dalloc(m), *u = dalloc(m);
  spmv(M_(c, FDAL_MAT_M), x2, t, 1.0, 0.0);
  apply_winv(c, t, u);
  spmv(M_(c, FDAL_MAT_CT), u, y0, -c->cfg.gamma, 1.0);
  free(t);
  free(u);
}
static void apply_a21_add(fdalo_ctx *c, const double *x0, double *y1) {
  const int64_t m = c->winv_n;
  double *t = dalloc(m), *u = dalloc(m);
  apply_C(c, x0, t);
  apply_winv(c, t, u);
  spmv(M_(c, FDAL_MAT_M), u, y1, -c->cfg.gamma2, 1.0);
  free(t);
  free(u);
}

/* AA / system_operator: block_operator row i = sum_j op(i,j) src_j, null blocks skipped
 *  */
static void apply_system(fdalo_ctx *c, const double *x, double *y) {
  const double *x0 = x, *x1 = x + c->n0, *x2 = x + c->n0 + c->n1;
  double *y0 = y, *y1 = y + c->n0, *y2 = y + c->n0 + c->n1;
  switch (c->cfg.kind) {
    case FDAL_KIND_LAPLACE:
      apply_aug(c, FDAL_AMG_A11, x0, y0);
      spmv(M_(c, FDAL_MAT_CT), x1, y0, 1.0, 1.0);
      apply_C(c, x0, y1);
      break;
    case FDAL_KIND_STOKES:
    case FDAL_KIND_STOKES_DIAG_MINRES:
      apply_aug(c, FDAL_AMG_A11, x0, y0);
      spmv(M_(c, FDAL_MAT_BT), x1, y0, 1.0, 1.0);
      spmv(M_(c, FDAL_MAT_CT), x2, y0, 1.0, 1.0);
      apply_B(c, x0, y1);
      apply_C(c, x0, y2);
      break;
    default: /* elliptic */
      apply_aug(c, FDAL_AMG_A11, x0, y0);
      apply_a12_add(c, x1, y0);
      spmv(M_(c, FDAL_MAT_CT), x2, y0, 1.0, 1.0);
      apply_aug(c, FDAL_AMG_A22, x1, y1);
      apply_a21_add(c, x0, y1);
      spmv(M_(c, FDAL_MAT_M), x2, y1, -1.0, 1.0);
      apply_C(c, x0, y2);
      spmv(M_(c, FDAL_MAT_M), x1, y2, -1.0, 1.0);
      break;
  }
}

/* inner preconditioner / operator adaptors for cg_solve */
static void prec_amg(fdalo_ctx *c, int which, const double *r, double *z) { amg_apply(&c->amg[which], r, z); }
static void op_aug(fdalo_ctx *c, int which, const double *x, double *y) { apply_aug(c, which, x, y); }
/* ideal elliptic: Aug 2x2 = [[A11g, A12g],[A21g, A22g]], prec = diag(AMG_A1, AMG_A2)
 * (elliptic_interface.cc:930-942) */
static void op_aug_block(fdalo_ctx *c, int a, const double *x, double *y) {
  (void)a;
  apply_aug(c, FDAL_AMG_A11, x, y);
  apply_a12_add(c, x + c->n0, y);
  apply_aug(c, FDAL_AMG_A22, x + c->n0, y + c->n0);
  apply_a21_add(c, x, y + c->n0);
}
static void prec_amg_block(fdalo_ctx *c, int a, const double *r, double *z) {
  (void)a;
  amg_apply(&c->amg[0], r, z);
  amg_apply(&c->amg[1], r + c->n0, z + c->n0);
}

/* Aug_inv = inverse_operator(Aug, SolverCG, AMG) */
static int apply_aug_inv(fdalo_ctx *c, int which, const double *b, double *x, int *its) {
  const int64_t n = which == FDAL_AMG_A11 ? c->n0 : c->n1;
  apply_fn pr = c->cfg.inner_prec == FDAL_PREC_AMG ? prec_amg : NULL;
  int st = cg_solve(c, n, op_aug, which, pr, which, &c->cfg.inner, b, x, its);
  if (which == FDAL_AMG_A11)
    c->its_a11 += *its;
  else
    c->its_a22 += *its;
  c->n_inner_solves++;
  if (st != FDAL_OK && !c->fail) c->fail = st;
  return st;
}

/* the five preconditioner vmults */
static void apply_prec(fdalo_ctx *c, const double *u, double *v) {
  const double *u0 = u, *u1 = u + c->n0, *u2 = u + c->n0 + c->n1;
  double *v0 = v, *v1 = v + c->n0, *v2 = v + c->n0 + c->n1;
  const double g = c->cfg.gamma;
  int its;
  switch (c->cfg.kind) {
    case FDAL_KIND_LAPLACE: {
      /* augmented_lagrangian_preconditioner.h:32-33 */
      apply_winv(c, u1, v1);
      vscale(c->n1, -g, v1);
      double *t = dalloc(c->n0);
      vcopy(c->n0, u0, t);
      spmv(M_(c, FDAL_MAT_CT), v1, t, -1.0, 1.0);
      apply_aug_inv(c, FDAL_AMG_A11, t, v0, &its);
      free(t);
    } break;
    case FDAL_KIND_STOKES: {
      /* :67-69 */
      apply_winv(c, u2, v2);
      vscale(c->n2, -g, v2);
      apply_mp_inv(c, u1, v1);
      vscale(c->n1, -c->cfg.gamma_grad_div, v1);
      double *t = dalloc(c->n0);
      vcopy(c->n0, u0, t);
      spmv(M_(c, FDAL_MAT_BT), v1, t, -1.0, 1.0);
      spmv(M_(c, FDAL_MAT_CT), v2, t, -1.0, 1.0);
      apply_aug_inv(c, FDAL_AMG_A11, t, v0, &its);
      free(t);
    } break;
    case FDAL_KIND_STOKES_DIAG_MINRES: {
      /* :100-102 */
      apply_winv(c, u2, v2);
      vscale(c->n2, g, v2);
      apply_mp_inv(c, u1, v1);
      vscale(c->n1, c->cfg.gamma_grad_div, v1);
      apply_aug_inv(c, FDAL_AMG_A11, u0, v0, &its);
    } break;
    case FDAL_KIND_ELLIPTIC_IDEAL: {
      /* :135-155 */
      apply_winv(c, u2, v2);
      vscale(c->n2, -g, v2);
      const int64_t nn = c->n0 + c->n1;
      double *uu = dalloc(nn);
      vcopy(c->n0, u0, uu);
      spmv(M_(c, FDAL_MAT_CT), v2, uu, -1.0, 1.0);
      vcopy(c->n1, u1, uu + c->n0);
      spmv(M_(c, FDAL_MAT_M), v2, uu + c->n0, 1.0, 1.0);
      int st = cg_solve(c, nn, op_aug_block, 0, c->cfg.inner_prec == FDAL_PREC_AMG ? prec_amg_block : NULL, 0,
                        &c->cfg.inner, uu, v, &its);
      c->its_a11 += its;
      c->n_inner_solves++;
      if (st != FDAL_OK && !c->fail) c->fail = st;
      free(uu);
    } break;
    case FDAL_KIND_ELLIPTIC_MODIFIED: {
      /* :225-228 */
      apply_winv(c, u2, v2);
      vscale(c->n2, -g, v2);
      double *t1 = dalloc(c->n1);
      vcopy(c->n1, u1, t1);
      spmv(M_(c, FDAL_MAT_M), v2, t1, 1.0, 1.0);
      apply_aug_inv(c, FDAL_AMG_A22, t1, v1, &its);
      /* u + gamma Ct invW M d1 - Ct d2 */
      double *t0 = dalloc(c->n0), *a = dalloc(c->n2), *b = dalloc(c->n2);
      vcopy(c->n0, u0, t0);
      spmv(M_(c, FDAL_MAT_M), v1, a, 1.0, 0.0);
      apply_winv(c, a, b);
      spmv(M_(c, FDAL_MAT_CT), b, t0, g, 1.0);
      spmv(M_(c, FDAL_MAT_CT), v2, t0, -1.0, 1.0);
      apply_aug_inv(c, FDAL_AMG_A11, t0, v0, &its);
      free(t1);
      free(t0);
      free(a);
      free(b);
    } break;
  }
}

/* ------------------------------------------------------------------ outer FGMRES */
static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static void record(fdal_solve_info *info, double res) {
  if (info->n_history < FDAL_MAX_HISTORY) info->residual_history[info->n_history++] = res;
}

static int fgmres(fdalo_ctx *c, const double *b, double *x, fdal_solve_info *info) {
  const int64_t N = c->N;
  const int mb = c->cfg.restart;
  double **V = (double **)malloc((size_t)(mb + 1) * sizeof(double *));
  double **Z = (double **)malloc((size_t)mb * sizeof(double *));
  for (int i = 0; i <= mb; ++i) V[i] = dalloc(N);
  for (int i = 0; i < mb; ++i) Z[i] = dalloc(N);
  double *H = dalloc((int64_t)(mb + 1) * mb); /* column-major, ld = mb+1 */
  double *g = dalloc(mb + 1), *cs = dalloc(mb), *sn = dalloc(mb), *y = dalloc(mb), *h2 = dalloc(mb + 1);
  control_state ctl;
  memset(&ctl, 0, sizeof(ctl));
  ctl.c = c->cfg.outer;
  int acc = 0, st = ST_ITERATE;
  do {
    /* v0 = b - A x */
    apply_system(c, x, V[0]);
    vsadd(N, -1.0, 1.0, b, V[0]);
    double res = sqrt(vdot(N, V[0], V[0]));
    if (acc == 0) info->initial_residual = res;
    st = control_check(&ctl, acc, res);
    if (acc == 0) record(info, res);
    if (st != ST_ITERATE) break;
    vscale(N, 1.0 / res, V[0]);
    memset(g, 0, (size_t)(mb + 1) * sizeof(double));
    g[0] = res;
    int j = 0;
    for (; j < mb && st == ST_ITERATE; ++j) {
      apply_prec(c, V[j], Z[j]);
      if (c->fail) break;
      double *w = V[j + 1];
      apply_system(c, Z[j], w);
      /* Orthogonalisation.  Default: MODIFIED Gram-Schmidt with deal.II's re-orthogonalisation rule
       * (LinearAlgebra::OrthogonalizationStrategy::modified_gram_schmidt: project out the basis
       * vectors one after the other; a second sweep only if the vector lost more than a factor
       * 10 sqrt(eps) of its length) — deliberately NOT the batched classical Gram-Schmidt with an
       * unconditional second pass that the CUDA library runs (csrc/fdal.cu: fgmres), so that the two
       * sides of every parity test orthogonalise with different algorithms; any numerically
       * orthonormal Arnoldi basis (deal.II >= 9.5 defaults to delayed classical Gram-Schmidt) gives
       * the same Hessenberg matrix to rounding.  FDALO_ORTHO=cgs2 selects the CUDA library's variant. */
      double *h = H + (int64_t)j * (mb + 1);
      static int ortho_mode = -1;
      if (ortho_mode < 0) {
        const char *e = getenv("FDALO_ORTHO");
        ortho_mode = (e && strcmp(e, "cgs2") == 0) ? 1 : 0;
      }
      if (ortho_mode == 1) {
        for (int i = 0; i <= j; ++i) h[i] = vdot(N, w, V[i]);
        for (int i = 0; i <= j; ++i) vaxpy(N, -h[i], V[i], w);
        for (int i = 0; i <= j; ++i) h2[i] = vdot(N, w, V[i]);
        for (int i = 0; i <= j; ++i) vaxpy(N, -h2[i], V[i], w);
        for (int i = 0; i <= j; ++i) h[i] += h2[i];
      } else {
        const double norm_start = sqrt(vdot(N, w, w));
        for (int i = 0; i <= j; ++i) {
          h[i] = vdot(N, w, V[i]);
          vaxpy(N, -h[i], V[i], w);
        }
        const double norm_after = sqrt(vdot(N, w, w));
        if (norm_after <= 10.0 * norm_start * sqrt(2.220446049250313e-16)) {
          for (int i = 0; i <= j; ++i) {
            const double t = vdot(N, w, V[i]);
            h[i] += t;
            vaxpy(N, -t, V[i], w);
          }
        }
      }
      const double hn = sqrt(vdot(N, w, w));
      h[j + 1] = hn;
      if (hn != 0.0) vscale(N, 1.0 / hn, w);
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * h[i] + sn[i] * h[i + 1];
        h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
        h[i] = t;
      }
      const double den = hypot(h[j], h[j + 1]);
      cs[j] = h[j] / den;
      sn[j] = h[j + 1] / den;
      h[j] = den;
      h[j + 1] = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      res = fabs(g[j + 1]);
      ++acc;
      record(info, res);
      st = control_check(&ctl, acc, res);
    }
    if (c->fail) {
      st = ST_FAILURE;
    }
    /* H y = g ; x += Z y */
    const int k = j;
    for (int i = k - 1; i >= 0; --i) {
      double s = g[i];
      for (int l = i + 1; l < k; ++l) s -= H[(int64_t)l * (mb + 1) + i] * y[l];
      y[i] = s / H[(int64_t)i * (mb + 1) + i];
    }
    for (int i = 0; i < k; ++i) vaxpy(N, y[i], Z[i], x);
    if (c->fail) break;
  } while (st == ST_ITERATE);
  info->outer_iterations = ctl.last_step;
  info->final_residual = ctl.last_value;
  for (int i = 0; i <= mb; ++i) free(V[i]);
  for (int i = 0; i < mb; ++i) free(Z[i]);
  free(V);
  free(Z);
  free(H);
  free(g);
  free(cs);
  free(sn);
  free(y);
  free(h2);
  if (c->fail) return c->fail;
  return st == ST_SUCCESS ? FDAL_OK : FDAL_ERR_OUTER_NO_CONVERGENCE;
}

/* ------------------------------------------------------------------ outer MinRes
 * deal.II SolverMinRes (stokes_immersed_boundary.cc:1057-1064): preconditioned
 * Lanczos with the residual measured in the preconditioner norm. */
static int minres(fdalo_ctx *c, const double *b, double *x, fdal_solve_info *info) {
  const int64_t N = c->N;
  double *u[3], *m[3], *v = dalloc(N);
  for (int i = 0; i < 3; ++i) {
    u[i] = dalloc(N);
    m[i] = dalloc(N);
  }
  double delta[3] = {0, 0, 0}, f[2] = {0, 0}, e[2] = {0, 0};
  double r_l2, r0, tau = 0, cc = 0, s = 0, d_ = 0;
  control_state ctl;
  memset(&ctl, 0, sizeof(ctl));
  ctl.c = c->cfg.outer;
  int j = 1;
  apply_system(c, x, m[0]);
  vcopy(N, b, u[1]);
  vaxpy(N, -1.0, m[0], u[1]);
  apply_prec(c, u[1], v);
  delta[1] = vdot(N, v, u[1]);
  r0 = sqrt(fabs(delta[1]));
  r_l2 = r0;
  vzero(N, u[0]);
  vzero(N, u[2]);
  vzero(N, m[0]);
  vzero(N, m[1]);
  vzero(N, m[2]);
  info->initial_residual = r_l2;
  record(info, r_l2);
  int st = control_check(&ctl, 0, r_l2);
  while (st == ST_ITERATE && !c->fail) {
    if (delta[1] != 0)
      vscale(N, 1.0 / sqrt(delta[1]), v);
    else
      vzero(N, v);
    apply_system(c, v, u[2]);
    if (j > 1) vaxpy(N, -sqrt(delta[1] / delta[0]), u[0], u[2]);
    const double gamma = vdot(N, u[2], v);
    vaxpy(N, -gamma / sqrt(delta[1]), u[1], u[2]);
    vcopy(N, v, m[0]);
    apply_prec(c, u[2], v);
    delta[2] = vdot(N, v, u[2]);
    if (j == 1) {
      d_ = gamma;
      e[1] = sqrt(fabs(delta[2]));
    }
    if (j > 1) {
      d_ = s * e[0] - cc * gamma;
      e[0] = cc * e[0] + s * gamma;
      f[1] = s * sqrt(fabs(delta[2]));
      e[1] = -cc * sqrt(fabs(delta[2]));
    }
    const double d = sqrt(d_ * d_ + fabs(delta[2]));
    if (j > 1) tau *= s / cc;
    cc = d_ / d;
    tau *= cc;
    s = sqrt(fabs(delta[2])) / d;
    if (j == 1) tau = r0 * cc;
    vaxpy(N, -e[0], m[1], m[0]);
    if (j > 1) vaxpy(N, -f[0], m[2], m[0]);
    vscale(N, 1.0 / d, m[0]);
    vaxpy(N, tau, m[0], x);
    r_l2 *= fabs(s);
    record(info, r_l2);
    st = control_check(&ctl, j, r_l2);
    ++j;
    double *t = m[2];
    m[2] = m[1];
    m[1] = m[0];
    m[0] = t;
    t = u[0];
    u[0] = u[1];
    u[1] = u[2];
    u[2] = t;
    delta[0] = delta[1];
    delta[1] = delta[2];
    f[0] = f[1];
    e[0] = e[1];
  }
  info->outer_iterations = ctl.last_step;
  info->final_residual = ctl.last_value;
  for (int i = 0; i < 3; ++i) {
    free(u[i]);
    free(m[i]);
  }
  free(v);
  if (c->fail) return c->fail;
  return st == ST_SUCCESS ? FDAL_OK : FDAL_ERR_OUTER_NO_CONVERGENCE;
}

/* ================================================================== C ABI (mirror of fdal.h) */
#define CHECK_CTX(c) \
  if (!(c)) return FDAL_ERR_INVALID
#define NEED_FINAL(c)                                   \
  if (!(c)->finalized) {                                \
    snprintf((c)->err, 512, "call fdalo_finalize first"); \
    return FDAL_ERR_STATE;                              \
  }

int fdalo_create(fdalo_ctx **out, const fdal_config *cfg) {
  if (!out || !cfg) return FDAL_ERR_INVALID;
  if (cfg->kind < 0 || cfg->kind > FDAL_KIND_ELLIPTIC_MODIFIED) return FDAL_ERR_INVALID;
  fdalo_ctx *c = (fdalo_ctx *)calloc(1, sizeof(fdalo_ctx));
  if (!c) return FDAL_ERR_ALLOC;
  c->cfg = *cfg;
  if (c->cfg.restart <= 0) c->cfg.restart = 30;
  *out = c;
  return FDAL_OK;
}
void fdalo_destroy(fdalo_ctx *c) {
  if (!c) return;
  for (int i = 0; i < FDAL_MAT_COUNT; ++i) csr_free(&c->mat[i]);
  free(c->winv_diag);
  free(c->mp_lumped_inv);
  lu_free(&c->lu_m);
  lu_free(&c->lu_mp);
  for (int a = 0; a < 2; ++a) {
    amg_t *g = &c->amg[a];
    for (int l = 0; l < g->nlev; ++l) {
      amg_level *L = &g->lev[l];
      csr_free(&L->A);
      csr_free(&L->P);
      csr_free(&L->R);
      free(L->inv_diag);
      free(L->x);
      free(L->b);
      free(L->r);
      free(L->d);
    }
    free(g->clu);
    free(g->cpiv);
  }
  free(c);
}
const char *fdalo_last_error(const fdalo_ctx *c) { return c ? c->err : "null context"; }
const char *fdalo_version(void) { return "fdal-oracle 0.1 (CPU restatement, test infrastructure)"; }

int fdalo_set_csr(fdalo_ctx *c, int id, int64_t nr, int64_t nc, int64_t nnz, const int64_t *rp,
                  const int32_t *ci, const double *v) {
  CHECK_CTX(c);
  if (id < 0 || id >= FDAL_MAT_COUNT || !rp || (nnz && (!ci || !v))) return FDAL_ERR_INVALID;
  if (rp[0] != 0 || rp[nr] != nnz) {
    snprintf(c->err, 512, "matrix %d: row_ptr inconsistent with nnz", id);
    return FDAL_ERR_SHAPE;
  }
  for (int64_t k = 0; k < nnz; ++k)
    if (ci[k] < 0 || ci[k] >= nc) {
      snprintf(c->err, 512, "matrix %d: column index out of range", id);
      return FDAL_ERR_SHAPE;
    }
  c->finalized = 0;
  return csr_copy(&c->mat[id], nr, nc, nnz, rp, ci, v);
}
int fdalo_set_diag(fdalo_ctx *c, int id, int64_t n, const double *d) {
  CHECK_CTX(c);
  if (!d || n < 0) return FDAL_ERR_INVALID;
  double **dst = id == FDAL_DIAG_W_INV ? &c->winv_diag : id == FDAL_DIAG_MP_LUMPED_INV ? &c->mp_lumped_inv : NULL;
  if (!dst) return FDAL_ERR_INVALID;
  free(*dst);
  *dst = dalloc(n);
  memcpy(*dst, d, (size_t)n * sizeof(double));
  if (id == FDAL_DIAG_W_INV)
    c->winv_n = n;
  else
    c->mp_n = n;
  c->finalized = 0;
  return FDAL_OK;
}
/* oracle-only: SuperLU factors of M (which=0) or Mp (which=1) — the stand-in
 * for SparseDirectUMFPACK::initialize (immersed_laplace.cc:644-645) */
int fdalo_set_lu(fdalo_ctx *c, int which, int64_t n, int64_t lnnz, const int64_t *Lp, const int32_t *Li,
                 const double *Lx, int64_t unnz, const int64_t *Up, const int32_t *Ui, const double *Ux,
                 const int32_t *perm_r, const int32_t *perm_c) {
  CHECK_CTX(c);
  lu_t *f = which == 0 ? &c->lu_m : &c->lu_mp;
  lu_free(f);
  f->n = n;
  f->Lp = (int64_t *)malloc((size_t)(n + 1) * 8);
  f->Up = (int64_t *)malloc((size_t)(n + 1) * 8);
  f->Li = (int32_t *)malloc((size_t)lnnz * 4);
  f->Ui = (int32_t *)malloc((size_t)unnz * 4);
  f->Lx = (double *)malloc((size_t)lnnz * 8);
  f->Ux = (double *)malloc((size_t)unnz * 8);
  f->perm_r = (int32_t *)malloc((size_t)n * 4);
  f->perm_c = (int32_t *)malloc((size_t)n * 4);
  f->w = dalloc(n);
  memcpy(f->Lp, Lp, (size_t)(n + 1) * 8);
  memcpy(f->Up, Up, (size_t)(n + 1) * 8);
  memcpy(f->Li, Li, (size_t)lnnz * 4);
  memcpy(f->Ui, Ui, (size_t)unnz * 4);
  memcpy(f->Lx, Lx, (size_t)lnnz * 8);
  memcpy(f->Ux, Ux, (size_t)unnz * 8);
  memcpy(f->perm_r, perm_r, (size_t)n * 4);
  memcpy(f->perm_c, perm_c, (size_t)n * 4);
  f->set = 1;
  return FDAL_OK;
}

int fdalo_amg_set_level(fdalo_ctx *c, int which, int level, const fdal_csr_view *A, const fdal_csr_view *P,
                        const fdal_csr_view *R, const double *inv_diag, double lmax, int degree, double ratio) {
  CHECK_CTX(c);
  if (which < 0 || which > 1 || level < 0 || level >= MAX_LEVELS || !A) return FDAL_ERR_INVALID;
  amg_t *g = &c->amg[which];
  amg_level *L = &g->lev[level];
  int st = csr_from_view(&L->A, A);
  if (st) return st;
  if (P) {
    if (P->n_rows != A->n_rows) {
      snprintf(c->err, 512, "AMG level %d: P has %ld rows, A has %ld", level, (long)P->n_rows, (long)A->n_rows);
      return FDAL_ERR_SHAPE;
    }
    st = csr_from_view(&L->P, P);
    if (st) return st;
  }
  if (R) {
    st = csr_from_view(&L->R, R);
    if (st) return st;
  }
  free(L->inv_diag);
  L->inv_diag = NULL;
  if (inv_diag) {
    L->inv_diag = dalloc(A->n_rows);
    memcpy(L->inv_diag, inv_diag, (size_t)A->n_rows * sizeof(double));
  }
  L->lmax = lmax;
  L->degree = degree;
  L->ratio = ratio;
  if (level + 1 > g->nlev) g->nlev = level + 1;
  c->finalized = 0;
  return FDAL_OK;
}
int fdalo_amg_set_coarse(fdalo_ctx *c, int which, int level, const fdal_csr_view *A) {
  return fdalo_amg_set_level(c, which, level, A, NULL, NULL, NULL, 1.0, 0, 1.0);
}

static int need(fdalo_ctx *c, int id, const char *name) {
  if (!c->mat[id].set) {
    snprintf(c->err, 512, "matrix %s not set", name);
    return 0;
  }
  return 1;
}
int fdalo_finalize(fdalo_ctx *c) {
  CHECK_CTX(c);
  const int k = c->cfg.kind;
  if (!need(c, FDAL_MAT_A, "A") || !need(c, FDAL_MAT_CT, "Ct")) return FDAL_ERR_STATE;
  c->n0 = c->mat[FDAL_MAT_A].nr;
  const int64_t m = c->mat[FDAL_MAT_CT].nc;
  if (c->mat[FDAL_MAT_CT].nr != c->n0 || c->mat[FDAL_MAT_A].nc != c->n0) {
    snprintf(c->err, 512, "A must be n x n and Ct n x m");
    return FDAL_ERR_SHAPE;
  }
  if (k == FDAL_KIND_LAPLACE) {
    c->nblocks = 2;
    c->n1 = m;
    c->n2 = 0;
  } else if (k == FDAL_KIND_STOKES || k == FDAL_KIND_STOKES_DIAG_MINRES) {
    if (!need(c, FDAL_MAT_BT, "Bt")) return FDAL_ERR_STATE;
    c->nblocks = 3;
    c->n1 = c->mat[FDAL_MAT_BT].nc;
    c->n2 = m;
    if (c->mat[FDAL_MAT_BT].nr != c->n0) return FDAL_ERR_SHAPE;
    if (c->cfg.mp_inv_mode == FDAL_MPINV_CG_LUMPED) {
      if (!need(c, FDAL_MAT_MP, "Mp")) return FDAL_ERR_STATE;
      if (!c->mp_lumped_inv) {
        /* M_p * 1, inverted (stokes_immersed_boundary.cc:946-952) */
        c->mp_n = c->n1;
        c->mp_lumped_inv = dalloc(c->n1);
        double *ones = dalloc(c->n1);
        for (int64_t i = 0; i < c->n1; ++i) ones[i] = 1.0;
        spmv(&c->mat[FDAL_MAT_MP], ones, c->mp_lumped_inv, 1.0, 0.0);
        for (int64_t i = 0; i < c->n1; ++i) c->mp_lumped_inv[i] = 1.0 / c->mp_lumped_inv[i];
        free(ones);
      }
    } else if (!c->lu_mp.set) {
      snprintf(c->err, 512, "exact Mp^-1 requested but no LU factors set");
      return FDAL_ERR_STATE;
    }
  } else {
    if (!need(c, FDAL_MAT_A2, "A2") || !need(c, FDAL_MAT_M, "M")) return FDAL_ERR_STATE;
    c->nblocks = 3;
    c->n1 = m;
    c->n2 = m;
    if (c->mat[FDAL_MAT_A2].nr != m || c->mat[FDAL_MAT_M].nr != m) return FDAL_ERR_SHAPE;
  }
  c->N = c->n0 + c->n1 + c->n2;
  if (c->cfg.winv_mode == FDAL_WINV_DIAG) {
    if (!c->winv_diag || c->winv_n != m) {
      snprintf(c->err, 512, "diagonal W^-1 of size m=%ld required", (long)m);
      return FDAL_ERR_STATE;
    }
  } else {
    if (!c->lu_m.set || c->lu_m.n != m) {
      snprintf(c->err, 512, "exact W^-1 requested but no LU factors of M set");
      return FDAL_ERR_STATE;
    }
    c->winv_n = m;
  }
  if (c->cfg.inner_prec == FDAL_PREC_AMG) {
    for (int a = 0; a < 2; ++a) {
      if (a == 1 && k != FDAL_KIND_ELLIPTIC_IDEAL && k != FDAL_KIND_ELLIPTIC_MODIFIED) continue;
      if (c->amg[a].nlev == 0) {
        snprintf(c->err, 512, "AMG hierarchy %d not set", a);
        return FDAL_ERR_STATE;
      }
      if (!c->amg[a].ready) {
        int st = amg_prepare(&c->amg[a], c->err);
        if (st) return st;
      }
    }
  }
  c->finalized = 1;
  return FDAL_OK;
}
int fdalo_block_sizes(const fdalo_ctx *c, int64_t sizes[3], int *nb) {
  CHECK_CTX(c);
  sizes[0] = c->n0;
  sizes[1] = c->n1;
  sizes[2] = c->n2;
  *nb = c->nblocks;
  return FDAL_OK;
}

int fdalo_spmv(fdalo_ctx *c, int id, int transpose, const double *x, double *y) {
  CHECK_CTX(c);
  if (id < 0 || id >= FDAL_MAT_COUNT || !c->mat[id].set) return FDAL_ERR_INVALID;
  if (transpose)
    spmv_t(&c->mat[id], x, y, 1.0, 0.0);
  else
    spmv(&c->mat[id], x, y, 1.0, 0.0);
  return FDAL_OK;
}
int fdalo_apply_aug(fdalo_ctx *c, int which, const double *x, double *y) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  c->fail = 0;
  apply_aug(c, which, x, y);
  return c->fail;
}
int fdalo_apply_system(fdalo_ctx *c, const double *x, double *y) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  c->fail = 0;
  apply_system(c, x, y);
  return c->fail;
}
int fdalo_apply_winv(fdalo_ctx *c, const double *x, double *y) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  apply_winv(c, x, y);
  return FDAL_OK;
}
int fdalo_apply_mp_inv(fdalo_ctx *c, const double *x, double *y, int *its) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  c->fail = 0;
  c->its_mass = 0;
  apply_mp_inv(c, x, y);
  if (its) *its = c->its_mass;
  return c->fail;
}
int fdalo_apply_amg(fdalo_ctx *c, int which, const double *r, double *z) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  if (!c->amg[which].ready) return FDAL_ERR_STATE;
  amg_apply(&c->amg[which], r, z);
  return FDAL_OK;
}
int fdalo_apply_aug_inv(fdalo_ctx *c, int which, const double *b, double *x, int *its) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  c->fail = 0;
  int i = 0;
  int st = apply_aug_inv(c, which, b, x, &i);
  if (its) *its = i;
  return st ? st : c->fail;
}
int fdalo_apply_prec(fdalo_ctx *c, const double *u, double *v, int inner_its[2]) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  c->fail = 0;
  c->its_a11 = c->its_a22 = 0;
  apply_prec(c, u, v);
  if (inner_its) {
    inner_its[0] = c->its_a11;
    inner_its[1] = c->its_a22;
  }
  return c->fail;
}
int fdalo_augment_rhs(fdalo_ctx *c, double *rhs) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  /* tmp = gamma * Ct * invW * embedded_rhs ; rhs0 += tmp */
  const int64_t m = c->winv_n;
  double *g = rhs + c->n0 + (c->nblocks == 3 ? c->n1 : 0);
  double *t = dalloc(m);
  apply_winv(c, g, t);
  spmv(&c->mat[FDAL_MAT_CT], t, rhs, c->cfg.gamma, 1.0);
  free(t);
  return FDAL_OK;
}
int fdalo_solve(fdalo_ctx *c, const double *rhs, double *x, fdal_solve_info *info) {
  CHECK_CTX(c);
  NEED_FINAL(c);
  fdal_solve_info local;
  if (!info) info = &local;
  memset(info, 0, sizeof(*info));
  c->fail = 0;
  c->its_a11 = c->its_a22 = c->its_mass = c->n_inner_solves = 0;
  const double t0 = now_ms();
  int st = c->cfg.kind == FDAL_KIND_STOKES_DIAG_MINRES ? minres(c, rhs, x, info) : fgmres(c, rhs, x, info);
  info->solve_ms = now_ms() - t0;
  info->status = st;
  info->inner_iterations = c->its_a11;
  info->inner_iterations_a22 = c->its_a22;
  info->inner_solves = c->n_inner_solves;
  info->mass_iterations = c->its_mass;
  if (st == FDAL_ERR_INNER_NO_CONVERGENCE)
    snprintf(c->err, 512, "inner CG did not converge (SolverControl::NoConvergence)");
  else if (st == FDAL_ERR_OUTER_NO_CONVERGENCE)
    snprintf(c->err, 512, "outer solver did not converge after %d steps", info->outer_iterations);
  return st;
}
