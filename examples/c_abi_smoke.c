/* Minimal C client of the drop-in boundary (include/fdal.h): links against libfdal.so without
 * any C++ / torch dependency, builds a 2x2 immersed-Laplace-type context from CSR arrays and,
 * when a GPU is present, runs one solve.  Without a GPU fdal_create fails loudly
 * (FDAL_ERR_CUDA) — there is no CPU fallback.  Build:
 *   gcc -std=c99 -Iinclude examples/c_abi_smoke.c -L<pkg>/csrc -lfdal -Wl,-rpath,<pkg>/csrc -o c_abi_smoke */
#include <stdio.h>
#include <stdlib.h>

#include "fdal.h"

int main(void) {
  printf("%s\n", fdal_version());
  fdal_config cfg = {0};
  cfg.kind = FDAL_KIND_LAPLACE;
  cfg.restart = 30;
  cfg.gamma = 10.0;
  cfg.winv_mode = FDAL_WINV_DIAG;
  cfg.inner_prec = FDAL_PREC_IDENTITY; /* no AMG hierarchy in this toy */
  cfg.outer.type = FDAL_CONTROL_REDUCTION;
  cfg.outer.max_steps = 200;
  cfg.outer.tol = 1e-10;
  cfg.outer.reduce = 1e-12;
  cfg.inner.type = FDAL_CONTROL_SOLVER;
  cfg.inner.max_steps = 200;
  cfg.inner.tol = 1e-8;
  fdal_ctx *ctx = NULL;
  int st = fdal_create(&ctx, &cfg);
  if (st != FDAL_OK) {
    printf("fdal_create -> %d (no CUDA device: expected on a CPU box)\n", st);
    return st == FDAL_ERR_CUDA ? 0 : 1;
  }
  /* A = tridiag(-1, 2, -1) (n = 8), Ct = one multiplier coupled to unknowns 3 and 4 */
  enum { n = 8, m = 1 };
  int64_t arp[n + 1], crp[n + 1];
  int32_t aci[3 * n], cci[2];
  double av[3 * n], cv[2] = {0.5, 0.5}, winv[m] = {4.0};
  int64_t k = 0, kc = 0;
  for (int i = 0; i < n; ++i) {
    arp[i] = k;
    crp[i] = kc;
    aci[k] = i; av[k++] = 2.0; /* diagonal first, like dealii::SparseMatrix */
    if (i > 0) { aci[k] = i - 1; av[k++] = -1.0; }
    if (i < n - 1) { aci[k] = i + 1; av[k++] = -1.0; }
    if (i == 3 || i == 4) cci[kc++] = 0;
  }
  arp[n] = k;
  crp[n] = kc;
  if ((st = fdal_set_csr(ctx, FDAL_MAT_A, n, n, k, arp, aci, av)) ||
      (st = fdal_set_csr(ctx, FDAL_MAT_CT, n, m, kc, crp, cci, cv)) ||
      (st = fdal_set_diag(ctx, FDAL_DIAG_W_INV, m, winv)) || (st = fdal_finalize(ctx))) {
    printf("setup failed: %d %s\n", st, fdal_last_error(ctx));
    return 1;
  }
  double rhs[n + m] = {1, 1, 1, 1, 1, 1, 1, 1, 0.25}, x[n + m] = {0};
  if ((st = fdal_augment_rhs(ctx, rhs))) return 1;
  fdal_solve_info info;
  st = fdal_solve(ctx, rhs, x, &info);
  printf("status %d, %d outer / %d inner iterations, residual %.3e, constraint 0.5(u3+u4) = %.6f\n", st,
         info.outer_iterations, info.inner_iterations, info.final_residual, 0.5 * (x[3] + x[4]));
  fdal_destroy(ctx);
  return st;
}
