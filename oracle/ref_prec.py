"""Handle on ``oracle/_ref/libref_prec.so``: the reference's OWN
``augmented_lagrangian_preconditioner.h`` compiled (unmodified, from /root/reference) against
the deal.II stand-in types of ``oracle/ref_harness/dealii_stub``.

TEST INFRASTRUCTURE — imported only by tests/ and tests/golden/generate_ref_prec.py.  It exists
to PIN the oracle: the block algebra of the five preconditioner ``vmult``s is executed by the
reference code itself, with every LinearOperator it is handed backed by a context's operator
applications.  /root/reference is absent on the GPU box, so the library is only ever built in
the development container; elsewhere ``available()`` is False and the committed golden vectors
(tests/golden/ref_prec_vectors.npz) stand in.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from fictitious_domain_al_preconditioners_b200 import _binding as b

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_prec.so")
REFERENCE_HEADER = "/root/reference/augmented_lagrangian_preconditioner.h"

OP_AUG_INV, OP_A22_INV, OP_AUG_INV_BLOCK, OP_C, OP_CT, OP_BT, OP_INVW, OP_MP_INV, OP_M = range(9)
OP_NAMES = ["Aug_inv", "A22_inv", "Aug_inv(block)", "C", "Ct", "Bt", "invW", "Mp_inv", "M"]

_OPFN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int64, C.POINTER(C.c_double), C.c_int64)
_lib = None


def build(force=False):
    """Compile the harness when the reference is mounted; returns the path or None."""
    if not os.path.exists(REFERENCE_HEADER):
        return LIB_PATH if os.path.exists(LIB_PATH) else None
    src = os.path.join(_HERE, "ref_harness", "ref_prec_harness.cc")
    stub = os.path.join(_HERE, "ref_harness", "dealii_stub", "deal.II", "lac", "stub_core.h")
    newest = max(os.path.getmtime(p) for p in (src, stub, REFERENCE_HEADER))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return LIB_PATH


def available() -> bool:
    try:
        return build() is not None
    except subprocess.CalledProcessError:
        return False


def _load():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libref_prec.so needs /root/reference (development container only)")
        _lib = C.CDLL(path)
        _lib.ref_prec_vmult.restype = C.c_int
        _lib.ref_prec_vmult.argtypes = [C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int64), _OPFN, C.c_void_p,
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib.ref_prec_source.restype = C.c_char_p
    return _lib


def context_operators(ctx):
    """The LinearOperators the reference applications hand to the preconditioner classes
    (immersed_laplace.cc:884-930, stokes_immersed_boundary.cc:923-1053, elliptic_interface.cc:
    800-900), realised with a finalized context's entry points."""
    sizes = ctx.sizes
    n0 = sizes[0]
    nl = sizes[-1]

    def block_aug_inv(x):
        # Aug_inv on [u, u2]: the preconditioner with a zero multiplier residual is exactly that
        # block solve (augmented_lagrangian_preconditioner.h:135-155 with u.block(2) = 0)
        v, _ = ctx.apply_prec(np.concatenate([x, np.zeros(nl)]))
        return v[: x.size]

    return {
        OP_AUG_INV: lambda x: ctx.apply_aug_inv(x, b.AMG_A11)[0],
        OP_A22_INV: lambda x: ctx.apply_aug_inv(x, b.AMG_A22)[0],
        OP_AUG_INV_BLOCK: block_aug_inv,
        OP_C: lambda x: ctx.spmv(b.MAT_CT, x, transpose=True, n_out=nl),
        OP_CT: lambda x: ctx.spmv(b.MAT_CT, x, n_out=n0),
        OP_BT: lambda x: ctx.spmv(b.MAT_BT, x, n_out=n0),
        OP_INVW: ctx.apply_winv,
        OP_MP_INV: lambda x: ctx.apply_mp_inv(x)[0],
        OP_M: lambda x: ctx.spmv(b.MAT_M, x, n_out=sizes[1]),
    }


def reference_vmult(kind, gamma, gamma_grad_div, sizes, operators, u, trace=None):
    """v = P.vmult(u) executed by the reference header; ``operators`` maps OP_* to a callable
    ndarray -> ndarray; ``trace`` (a list) receives the OP_* ids in call order."""
    lib = _load()
    u = np.ascontiguousarray(u, dtype=np.float64)
    v = np.zeros_like(u)
    failure = []

    def cb(_user, op, x, nx, y, ny):
        try:
            if trace is not None:
                trace.append(op)
            out = operators[op](np.ctypeslib.as_array(x, shape=(nx,)).copy())
            assert out.shape == (ny,), (OP_NAMES[op], out.shape, ny)
            np.ctypeslib.as_array(y, shape=(ny,))[:] = out
        except BaseException as e:  # never let an exception cross the C frame
            failure.append(e)

    kind_id = {b.KIND_LAPLACE: 0, b.KIND_STOKES: 1, b.KIND_STOKES_DIAG_MINRES: 2, b.KIND_ELLIPTIC_IDEAL: 3,
               b.KIND_ELLIPTIC_MODIFIED: 4}[kind]
    sz = (C.c_int64 * 3)(*(list(sizes) + [0] * (3 - len(sizes))))
    st = lib.ref_prec_vmult(kind_id, float(gamma), float(gamma_grad_div), sz, _OPFN(cb), None, b.dptr(u), b.dptr(v))
    if failure:
        raise failure[0]
    if st != 0:
        raise RuntimeError(f"reference vmult failed with status {st}")
    return v
