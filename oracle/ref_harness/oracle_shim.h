/* Test infrastructure: lets include/fdal_dealii.h (which calls fdal_*) bind to the CPU oracle
 * (which exports the same ABI as fdalo_*), so the adapter can be exercised without a GPU. */
#ifndef FDAL_ORACLE_SHIM_H
#define FDAL_ORACLE_SHIM_H
#define fdal_last_error fdalo_last_error
#define fdal_set_csr fdalo_set_csr
#define fdal_amg_set_level fdalo_amg_set_level
#define fdal_amg_set_coarse fdalo_amg_set_coarse
#define fdal_spmv fdalo_spmv
#define fdal_apply_aug fdalo_apply_aug
#define fdal_apply_winv fdalo_apply_winv
#define fdal_apply_mp_inv fdalo_apply_mp_inv
#define fdal_apply_aug_inv fdalo_apply_aug_inv
#define fdal_apply_prec fdalo_apply_prec
#define fdal_solve fdalo_solve
#endif
