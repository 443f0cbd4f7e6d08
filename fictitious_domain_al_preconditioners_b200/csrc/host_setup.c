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
