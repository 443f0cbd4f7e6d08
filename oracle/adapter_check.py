"""Handle on ``oracle/_ref/libadapter_{oracle,cuda}.so`` (ref_harness/adapter_check.cc): the
reference-side binding ``include/fdal_dealii.h`` compiled against stand-in deal.II types together
with the reference's own preconditioner classes, bound to the CPU oracle or to the CUDA library.

TEST INFRASTRUCTURE — imported only by tests/.  Built only where /root/reference is mounted
(``make -C oracle adapter``); the prebuilt libraries travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fictitious_domain_al_preconditioners_b200 import _binding as b

from .ref_prec import REFERENCE_HEADER

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs: dict = {}


def lib_path(flavour: str) -> str:
    return os.path.join(_HERE, "_ref", f"libadapter_{flavour}.so")


def build(flavour: str):
    """Returns the library path, (re)building it when the reference is mounted; None if unavailable."""
    path = lib_path(flavour)
    if os.path.exists(REFERENCE_HEADER):
        try:
            subprocess.run(["make", "-C", _HERE, f"_ref/libadapter_{flavour}.so"], check=True, capture_output=True)
        except subprocess.CalledProcessError:
            return None
    return path if os.path.exists(path) else None


def available(flavour: str) -> bool:
    return build(flavour) is not None


def load(flavour: str):
    if flavour not in _libs:
        path = build(flavour)
        if path is None:
            raise RuntimeError(f"{lib_path(flavour)} needs /root/reference (development container only)")
        lib = C.CDLL(path)
        pd, pi64 = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        lib.adapter_reference_vmult.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, pi64, pd, pd]
        lib.adapter_al_vmult.argtypes = [C.c_void_p, C.c_int, pi64, pd, pd, C.POINTER(C.c_int)]
        lib.adapter_solve.argtypes = [C.c_void_p, C.c_int, pi64, pd, pd, C.POINTER(b.SolveInfo)]
        lib.adapter_export_csr.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, pi64, C.POINTER(C.c_int32), pd]
        lib.adapter_to_control.argtypes = [C.c_int, C.c_uint, C.c_double, C.c_double, C.POINTER(b.Control)]
        lib.adapter_to_control.restype = None
        pp32, ppd = C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_double))
        lib.adapter_export_amg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), pp32,
                                           pp32, ppd, pd, C.c_int, C.c_double]
        lib.adapter_export_amg.restype = C.c_int
        for f in ("adapter_reference_vmult", "adapter_al_vmult", "adapter_solve", "adapter_export_csr"):
            getattr(lib, f).restype = C.c_int
        _libs[flavour] = lib
    return _libs[flavour]


def _sizes(ctx):
    return (C.c_int64 * 3)(*(list(ctx.sizes) + [0] * (3 - len(ctx.sizes))))


def reference_vmult(lib, ctx, u):
    """The REFERENCE preconditioner class built from the adapter's LinearOperators on `ctx`."""
    cfg = ctx.config
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros_like(u)
    st = lib.adapter_reference_vmult(ctx._h, cfg.kind, cfg.gamma, cfg.gamma_grad_div, _sizes(ctx), b.dptr(u), b.dptr(v))
    return v, st


def al_vmult(lib, ctx, u):
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros_like(u)
    its = (C.c_int * 2)()
    st = lib.adapter_al_vmult(ctx._h, len(ctx.sizes), _sizes(ctx), b.dptr(u), b.dptr(v), its)
    return v, (its[0], its[1]), st


def solve(lib, ctx, rhs, x0=None):
    rhs = np.ascontiguousarray(rhs, dtype=np.float64)
    x = np.zeros_like(rhs) if x0 is None else np.ascontiguousarray(x0, dtype=np.float64).copy()
    info = b.SolveInfo()
    st = lib.adapter_solve(ctx._h, len(ctx.sizes), _sizes(ctx), b.dptr(rhs), b.dptr(x), C.byref(info))
    return x, info, st


def export_csr(lib, ctx, matrix_id, A):
    """fdal_dealii::export_csr from a dealii::SparseMatrix holding A; records it on the Python side too."""
    rp, ci, val = b.csr_arrays(A)
    st = lib.adapter_export_csr(ctx._h, matrix_id, A.shape[0], A.shape[1], rp.ctypes.data_as(C.POINTER(C.c_int64)),
                                ci.ctypes.data_as(C.POINTER(C.c_int32)), b.dptr(val))
    if st == 0:
        ctx.matrices.add(matrix_id)
    return st


def to_control(lib, type_, max_steps, tol, reduce=0.0):
    out = b.Control()
    lib.adapter_to_control(type_, max_steps, tol, reduce, C.byref(out))
    return out


def export_amg(lib, ctx, which, hierarchy):
    """fdal_dealii::export_amg from a TrilinosWrappers::PreconditionAMG stand-in whose ML hierarchy
    (ref_harness/trilinos_stub: Amat[l], Pmat[l+1], Rmat[l], Amat[l].lambda_max) is filled from `hierarchy`."""
    import scipy.sparse as sp

    levels = hierarchy.levels
    nl = len(levels)
    mats = []
    for l, L in enumerate(levels):
        mats.append(sp.csr_matrix(L.A))
        if l + 1 < nl:
            mats.append(sp.csr_matrix(L.P))
            mats.append(sp.csr_matrix(L.R if L.R is not None else L.P.T))
        else:
            mats += [None, None]
    keep, rp, ci, val = [], [], [], []
    n_rows = (C.c_int32 * (3 * nl))()
    n_cols = (C.c_int32 * (3 * nl))()
    for k, M in enumerate(mats):
        if M is None:
            a = (np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1))
        else:
            a = (np.ascontiguousarray(M.indptr, np.int32), np.ascontiguousarray(M.indices, np.int32),
                 np.ascontiguousarray(M.data, np.float64))
            n_rows[k], n_cols[k] = M.shape
        keep.append(a)
        rp.append(a[0].ctypes.data_as(C.POINTER(C.c_int32)))
        ci.append(a[1].ctypes.data_as(C.POINTER(C.c_int32)))
        val.append(a[2].ctypes.data_as(C.POINTER(C.c_double)))
    lam = np.array([L.lambda_max for L in levels], dtype=np.float64)
    return lib.adapter_export_amg(ctx._h, which, nl, n_rows, n_cols, (C.POINTER(C.c_int32) * len(rp))(*rp),
                                  (C.POINTER(C.c_int32) * len(ci))(*ci), (C.POINTER(C.c_double) * len(val))(*val),
                                  b.dptr(lam), int(hierarchy.cheb_degree), float(hierarchy.eig_ratio))
