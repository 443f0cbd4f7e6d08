"""Python handle on the CPU oracle (``oracle/libfdal_oracle.so``).

TEST INFRASTRUCTURE.  Imported only by tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs.  It reuses the
product's ctypes table (same ABI, prefix ``fdalo_``) so both sides of a parity
test are driven by identical calls.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200.context import ALConfig, ALContext

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfdal_oracle.so")
_api = None


def build(force=False):
    src = os.path.join(_HERE, "fdal_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "fdal.h")
    newest = max(os.path.getmtime(src), os.path.getmtime(hdr))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB_PATH


def load() -> b.Api:
    global _api
    if _api is None:
        build()
        _api = b.Api(LIB_PATH, "fdalo_", extra=b.ORACLE_SIGNATURES)
    return _api


class OracleContext(ALContext):
    """ALContext bound to the oracle; adds the SuperLU hand-over for exact mass inverses."""

    def __init__(self, config: ALConfig, threads: int = 1):
        api = load()
        api.set_num_threads(threads)
        super().__init__(config, api=api)

    def set_lu(self, which: int, M):
        """which: 0 = immersed mass M, 1 = pressure mass Mp (UMFPACK stand-in)."""
        lu = spla.splu(sp.csc_matrix(M))
        L, U = lu.L.tocsc(), lu.U.tocsc()
        n = M.shape[0]
        arrs = dict(
            Lp=np.ascontiguousarray(L.indptr, dtype=np.int64),
            Li=np.ascontiguousarray(L.indices, dtype=np.int32),
            Lx=np.ascontiguousarray(L.data, dtype=np.float64),
            Up=np.ascontiguousarray(U.indptr, dtype=np.int64),
            Ui=np.ascontiguousarray(U.indices, dtype=np.int32),
            Ux=np.ascontiguousarray(U.data, dtype=np.float64),
            pr=np.ascontiguousarray(lu.perm_r, dtype=np.int32),
            pc=np.ascontiguousarray(lu.perm_c, dtype=np.int32),
        )
        i64, i32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        self._check(
            self.api.set_lu(
                self._h, which, n,
                arrs["Lx"].size, arrs["Lp"].ctypes.data_as(i64), arrs["Li"].ctypes.data_as(i32), b.dptr(arrs["Lx"]),
                arrs["Ux"].size, arrs["Up"].ctypes.data_as(i64), arrs["Ui"].ctypes.data_as(i32), b.dptr(arrs["Ux"]),
                arrs["pr"].ctypes.data_as(i32), arrs["pc"].ctypes.data_as(i32),
            )
        )
