"""ctypes view of the C ABI declared in ``include/fdal.h``.

The same table binds the CUDA library (prefix ``fdal_``) and — from ``oracle/``
and the tests only — the CPU oracle (prefix ``fdalo_``); nothing in this
package ever loads the oracle.
"""
from __future__ import annotations

import ctypes as C
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

# ---- enums (keep in sync with include/fdal.h) --------------------------------
OK = 0
ERR_INVALID, ERR_SHAPE, ERR_ALLOC, ERR_CUDA, ERR_STATE = 1, 2, 3, 4, 5
ERR_INNER_NO_CONVERGENCE, ERR_OUTER_NO_CONVERGENCE, ERR_MASS_NO_CONVERGENCE = 6, 7, 8
ERR_NCCL, ERR_UNSUPPORTED = 9, 10

KIND_LAPLACE, KIND_STOKES, KIND_STOKES_DIAG_MINRES = 0, 1, 2
KIND_ELLIPTIC_IDEAL, KIND_ELLIPTIC_MODIFIED = 3, 4

MAT_A, MAT_A2, MAT_BT, MAT_B, MAT_CT, MAT_C, MAT_M, MAT_MP = range(8)
MAT_AMG_A, MAT_AMG_P, MAT_AMG_R = 100, 101, 102
WINV_DIAG, WINV_EXACT_M, WINV_EXACT_M_SQUARED = 0, 1, 2
MPINV_CG_LUMPED, MPINV_EXACT = 0, 1
DIAG_W_INV, DIAG_MP_LUMPED_INV = 0, 1
PREC_AMG, PREC_IDENTITY = 0, 1
AMG_A11, AMG_A22 = 0, 1
CONTROL_SOLVER, CONTROL_REDUCTION, CONTROL_ITERATION_NUMBER = 0, 1, 2

TIME_SPMV_A, TIME_AUG, TIME_VCYCLE, TIME_CHEB_FINE, TIME_DOT, TIME_MULTIDOT, TIME_AXPY = range(7)

MAX_HISTORY = 1024


class Control(C.Structure):
    _fields_ = [("type", c_int32), ("max_steps", c_int32), ("tol", c_double), ("reduce", c_double)]


class Config(C.Structure):
    _fields_ = [
        ("kind", c_int32),
        ("restart", c_int32),
        ("gamma", c_double),
        ("gamma2", c_double),
        ("gamma_grad_div", c_double),
        ("winv_mode", c_int32),
        ("mp_inv_mode", c_int32),
        ("aug_explicit", c_int32),
        ("grad_div_in_operator", c_int32),
        ("inner_prec", c_int32),
        ("device", c_int32),
        ("use_graphs", c_int32),
        ("exact_mass_max_its", c_int32),
        ("block_size", c_int32),
        ("reserved0", c_int32),
        ("outer", Control),
        ("inner", Control),
        ("mass", Control),
    ]


class SolveInfo(C.Structure):
    _fields_ = [
        ("status", c_int32),
        ("outer_iterations", c_int32),
        ("inner_iterations", c_int32),
        ("inner_iterations_a22", c_int32),
        ("inner_solves", c_int32),
        ("mass_iterations", c_int32),
        ("n_history", c_int32),
        ("reserved", c_int32),
        ("initial_residual", c_double),
        ("final_residual", c_double),
        ("solve_ms", c_double),
        ("kernel_launches", c_int64),
        ("residual_history", c_double * MAX_HISTORY),
    ]

    def history(self):
        return np.array(self.residual_history[: self.n_history])


class CsrView(C.Structure):
    _fields_ = [
        ("n_rows", c_int64),
        ("n_cols", c_int64),
        ("nnz", c_int64),
        ("row_ptr", POINTER(c_int64)),
        ("col", POINTER(c_int32)),
        ("val", POINTER(c_double)),
    ]


_pd = POINTER(c_double)
_pi32 = POINTER(c_int32)
_pi64 = POINTER(c_int64)
_pv = POINTER(CsrView)

# name -> (restype, argtypes); every symbol include/fdal.h declares
SIGNATURES = {
    "create": (c_int, [POINTER(c_void_p), POINTER(Config)]),
    "destroy": (None, [c_void_p]),
    "last_error": (c_char_p, [c_void_p]),
    "version": (c_char_p, []),
    "set_csr": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, _pi64, _pi32, _pd]),
    "set_diag": (c_int, [c_void_p, c_int, c_int64, _pd]),
    "amg_set_level": (c_int, [c_void_p, c_int, c_int, _pv, _pv, _pv, _pd, c_double, c_int, c_double]),
    "amg_set_coarse": (c_int, [c_void_p, c_int, c_int, _pv]),
    "finalize": (c_int, [c_void_p]),
    "block_sizes": (c_int, [c_void_p, POINTER(c_int64 * 3), POINTER(c_int)]),
    "spmv": (c_int, [c_void_p, c_int, c_int, _pd, _pd]),
    "apply_aug": (c_int, [c_void_p, c_int, _pd, _pd]),
    "apply_system": (c_int, [c_void_p, _pd, _pd]),
    "apply_winv": (c_int, [c_void_p, _pd, _pd]),
    "apply_mp_inv": (c_int, [c_void_p, _pd, _pd, POINTER(c_int)]),
    "apply_amg": (c_int, [c_void_p, c_int, _pd, _pd]),
    "apply_aug_inv": (c_int, [c_void_p, c_int, _pd, _pd, POINTER(c_int)]),
    "apply_prec": (c_int, [c_void_p, _pd, _pd, POINTER(c_int * 2)]),
    "augment_rhs": (c_int, [c_void_p, _pd]),
    "solve": (c_int, [c_void_p, _pd, _pd, POINTER(SolveInfo)]),
}
# only the CUDA library has these
DEVICE_SIGNATURES = {
    "solve_dev": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(SolveInfo)]),
    "apply_aug_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "apply_amg_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "spmv_dev": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "time_kernel": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int, c_int, _pd, _pd, _pi64],
    ),
    "nccl_unique_id": (c_int, [C.c_char * 128]),
    "comm_init": (c_int, [c_void_p, C.c_char * 128, c_int, c_int]),
    "set_halo": (
        c_int,
        [c_void_p, c_int, c_int, c_int, c_int64, c_int64, _pi32, _pi32, _pi32],
    ),
    "amg_set_coarse_range": (c_int, [c_void_p, c_int, c_int64, c_int64]),
    "amg_set_replicated_from": (c_int, [c_void_p, c_int, c_int]),
    "comm_mode": (c_int, [c_void_p]),
    "mass_solver_info": (c_int, [c_void_p, c_int, _pi32, _pi32, _pd, _pd, _pd]),
    "bsr_conversions": (c_int, [c_void_p, _pi32, _pi32]),
    "csr_to_bsr": (c_int64, [c_int, c_int64, c_int64, _pi64, _pi32, _pd, c_int32, c_double, _pi32, c_int64, _pi32, _pd]),
    "assemble_al_term": (c_int, [c_int, c_int64, _pi64, _pi32, _pd, c_int64, c_int32, _pi32, _pd, _pd, _pi64]),
}
# only the oracle has these
ORACLE_SIGNATURES = {
    "set_lu": (
        c_int,
        [c_void_p, c_int, c_int64, c_int64, _pi64, _pi32, _pd, c_int64, _pi64, _pi32, _pd, _pi32, _pi32],
    ),
    "set_num_threads": (None, [c_int]),
    "get_max_threads": (c_int, []),
}


class Api:
    """Function table of one shared library implementing the fdal ABI."""

    def __init__(self, path: str, prefix: str, extra: dict | None = None):
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "fdal_" else C.DEFAULT_MODE)
        table = dict(SIGNATURES)
        if extra:
            table.update(extra)
        for name, (res, args) in table.items():
            fn = getattr(self.lib, prefix + name)  # AttributeError if a symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_pd)


def csr_arrays(A):
    """(row_ptr int64, col int32, val f64) of a scipy CSR matrix, order preserved."""
    A = A.tocsr()
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    v = np.ascontiguousarray(A.data, dtype=np.float64)
    return rp, ci, v


def csr_view(A):
    """CsrView + the arrays that must stay alive while it is used."""
    rp, ci, v = csr_arrays(A)
    view = CsrView(
        A.shape[0],
        A.shape[1],
        v.size,
        rp.ctypes.data_as(_pi64),
        ci.ctypes.data_as(_pi32),
        v.ctypes.data_as(_pd),
    )
    return view, (rp, ci, v)
