/*
 * host_setup.c — host-side helpers for the stand-in AMG setup (setup phase,
 * once per solve; NOT on the per-iteration path, which is CUDA only).
 *
 * In the reference the hierarchy comes from Trilinos ML through
 * TrilinosWrappers::PreconditionAMG::initialize (utilities.h:308-317,
 * immersed_laplace.cc:833).  ML is not available in this image, so
 * amg_setup.py builds a smoothed-aggregation hierarchy of the same shape and
 * hands it over through fdal_amg_set_level exactly as an ML export would be.
 */
#include <stdint.h>
#include <stdlib.h>

/* Greedy three-pass aggregation on a symmetric strength graph (no self loops).
 * Returns the number of aggregates; agg[i] in [0, n_agg), or -1 for rows without
 * any strong connection (Dirichlet / constrained rows): like ML these are left
 * out of the coarse grid — the smoother alone treats them. */
int64_t fdal_host_aggregate(int64_t n, const int64_t *indptr, const int32_t *indices, int32_t *agg) {
  int64_t n_agg = 0;
  for (int64_t i = 0; i < n; ++i) agg[i] = -1;
  /* pass 1: a node whose whole strong neighbourhood is free seeds an aggregate */
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    int free_nb = 1, has_nb = 0;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      const int32_t j = indices[k];
      if (j == i) continue;
      has_nb = 1;
      if (agg[j] != -1) {
        free_nb = 0;
        break;
      }
    }
    if (!has_nb) continue; /* isolated (e.g. Dirichlet row) */
    if (!free_nb) continue;
    agg[i] = (int32_t)n_agg;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) agg[indices[k]] = (int32_t)n_agg;
    ++n_agg;
  }
  /* pass 2: attach leftovers to a neighbouring pass-1 aggregate */
  int32_t *tmp = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
  for (int64_t i = 0; i < n; ++i) tmp[i] = agg[i];
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      const int32_t j = indices[k];
      if (j != i && tmp[j] != -1) {
        agg[i] = tmp[j];
        break;
      }
    }
  }
  free(tmp);
  /* pass 3: whatever is left forms new aggregates with its free neighbours */
  for (int64_t i = 0; i < n; ++i) {
    if (agg[i] != -1) continue;
    if (indptr[i + 1] == indptr[i]) continue; /* isolated: not aggregated */
    agg[i] = (int32_t)n_agg;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) {
      const int32_t j = indices[k];
      if (agg[j] == -1) agg[j] = (int32_t)n_agg;
    }
    ++n_agg;
  }
  return n_agg;
}

/* ---- row-parallel Gustavson SpGEMM (OpenMP): C = A * B, all CSR ------------------------
 * scipy's csr_matmat is single-threaded and needs minutes for the Galerkin products of a
 * 10^9 non-zero fine level; this is the same algorithm with one dense accumulator per
 * thread.  Two passes: symbolic (row counts -> Cp), numeric (sorted columns + values). */
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int cmp_i32(const void *a, const void *b) {
  const int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return (x > y) - (x < y);
}

int64_t fdal_host_spgemm_symbolic(int64_t n_rows, int64_t n_cols_b, const int64_t *Ap, const int32_t *Aj,
                                  const int64_t *Bp, const int32_t *Bj, int64_t *Cp) {
  Cp[0] = 0;
#pragma omp parallel
  {
    int64_t *marker = (int64_t *)malloc((size_t)(n_cols_b > 0 ? n_cols_b : 1) * sizeof(int64_t));
    for (int64_t j = 0; j < n_cols_b; ++j) marker[j] = -1;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_rows; ++i) {
      int64_t cnt = 0;
      for (int64_t ka = Ap[i]; ka < Ap[i + 1]; ++ka) {
        const int32_t k = Aj[ka];
        for (int64_t kb = Bp[k]; kb < Bp[k + 1]; ++kb) {
          const int32_t j = Bj[kb];
          if (marker[j] != i) {
            marker[j] = i;
            ++cnt;
          }
        }
      }
      Cp[i + 1] = cnt;
    }
    free(marker);
  }
  for (int64_t i = 0; i < n_rows; ++i) Cp[i + 1] += Cp[i];
  return Cp[n_rows];
}

void fdal_host_spgemm_numeric(int64_t n_rows, int64_t n_cols_b, const int64_t *Ap, const int32_t *Aj,
                              const double *Ax, const int64_t *Bp, const int32_t *Bj, const double *Bx,
                              const int64_t *Cp, int32_t *Cj, double *Cx) {
#pragma omp parallel
  {
    int64_t *marker = (int64_t *)malloc((size_t)(n_cols_b > 0 ? n_cols_b : 1) * sizeof(int64_t));
    double *acc = (double *)malloc((size_t)(n_cols_b > 0 ? n_cols_b : 1) * sizeof(double));
    for (int64_t j = 0; j < n_cols_b; ++j) marker[j] = -1;
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_rows; ++i) {
      int32_t *cols = Cj + Cp[i];
      int64_t cnt = 0;
      for (int64_t ka = Ap[i]; ka < Ap[i + 1]; ++ka) {
        const int32_t k = Aj[ka];
        const double a = Ax[ka];
        for (int64_t kb = Bp[k]; kb < Bp[k + 1]; ++kb) {
          const int32_t j = Bj[kb];
          if (marker[j] != i) {
            marker[j] = i;
            acc[j] = a * Bx[kb];
            cols[cnt++] = j;
          } else {
            acc[j] += a * Bx[kb];
          }
        }
      }
      qsort(cols, (size_t)cnt, sizeof(int32_t), cmp_i32);
      double *vals = Cx + Cp[i];
      for (int64_t q = 0; q < cnt; ++q) vals[q] = acc[cols[q]];
    }
    free(marker);
    free(acc);
  }
}

/* thread budget of the OpenMP helpers (bench: rank 0 uses the whole box during the shared
 * setup, every rank its share afterwards); the CUDA library's BSR conversion uses the same
 * libgomp runtime */
void fdal_host_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n < 1 ? 1 : n);
#else
  (void)n;
#endif
}
int fdal_host_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- row-partition helpers (partition.py; rank-0 setup of the multi-GPU runs) ------------------- */
/* marks[c] = 1 for every column of indices[0:nnz) outside [c0, c1) (racy stores of the same value) */
void fdal_host_mark_foreign_cols(int64_t nnz, const int32_t *indices, int64_t c0, int64_t c1, uint8_t *marks) {
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t c = indices[k];
    if (c < c0 || c >= c1) marks[c] = 1;
  }
}
/* local column numbering [owned | halo]: owned c -> c - c0, foreign c -> n_owned + position of c in the
 * sorted list halo[0:n_halo) */
void fdal_host_localize_cols(int64_t nnz, const int32_t *indices, int64_t c0, int64_t c1, const int64_t *halo,
                             int64_t n_halo, int32_t *out) {
  const int64_t n_owned = c1 - c0;
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < nnz; ++k) {
    const int64_t c = indices[k];
    if (c >= c0 && c < c1) {
      out[k] = (int32_t)(c - c0);
    } else {
      int64_t lo = 0, hi = n_halo;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (halo[mid] < c)
          lo = mid + 1;
        else
          hi = mid;
      }
      out[k] = (int32_t)(n_owned + lo);
    }
  }
}
