"""Stand-in for ``TrilinosWrappers::PreconditionAMG::initialize`` (host, setup phase).

The reference builds its AMG with Trilinos ML on the explicit matrix
``A + gamma * Ct * diag(W^-1) * C`` (utilities.h:112-331, 591-744;
immersed_laplace.cc:709-846) and only *applies* it on the hot path.  ML is not
installed here, so this module builds a smoothed-aggregation hierarchy with the
same ingredients (symmetric strength-of-connection with a drop threshold,
greedy aggregation, piecewise-constant tentative prolongator on the constant
modes, damped-Jacobi prolongator smoothing with omega = 4/3 / lambda_max,
Galerkin RAP, coarsening until <= ``max_coarse`` rows) and exports it in the
hierarchy exchange format of ``fdal_amg_set_level`` — the format an adapter
would fill from ``ML_Epetra::MultiLevelPreconditioner``.

Nothing here runs per iteration; the V-cycle itself is CUDA (csrc/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST_LIB = os.path.join(_HERE, "csrc", "libfdal_host.so")
_lib = None


def build_host_lib(force=False):
    csrc = os.path.join(_HERE, "csrc")
    src = os.path.join(csrc, "host_setup.c")
    # the host half of fdal_finalize (csrc/host_finalize.h) rides along so the CPU tests can call it
    hooks = os.path.join(csrc, "host_finalize_hooks.cpp")
    deps = [src, hooks, os.path.join(csrc, "host_finalize.h")]
    if force or not os.path.exists(_HOST_LIB) or any(os.path.getmtime(_HOST_LIB) < os.path.getmtime(d) for d in deps):
        obj = os.path.join(csrc, "host_setup.o")
        subprocess.run(["/usr/bin/gcc", "-O3", "-fopenmp", "-fPIC", "-c", "-o", obj, src], check=True, capture_output=True)
        subprocess.run(
            ["/usr/bin/g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", _HOST_LIB, hooks, obj], check=True,
            capture_output=True
        )
    return _HOST_LIB


def _host():
    global _lib
    if _lib is None:
        build_host_lib()
        _lib = C.CDLL(_HOST_LIB)
        _lib.fdal_host_aggregate.restype = C.c_int64
        _lib.fdal_host_aggregate.argtypes = [
            C.c_int64,
            C.POINTER(C.c_int64),
            C.POINTER(C.c_int32),
            C.POINTER(C.c_int32),
        ]
        p64, p32, pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
        _lib.fdal_host_spgemm_symbolic.restype = C.c_int64
        _lib.fdal_host_spgemm_symbolic.argtypes = [C.c_int64, C.c_int64, p64, p32, p64, p32, p64]
        _lib.fdal_host_spgemm_numeric.restype = None
        _lib.fdal_host_spgemm_numeric.argtypes = [C.c_int64, C.c_int64, p64, p32, pd, p64, p32, pd, p64, p32, pd]
    return _lib


_GPU_SPGEMM_MIN_NNZ = 20_000_000


import sys
import time

_VERBOSE = bool(os.environ.get("FDAL_VERBOSE_SETUP"))


def set_host_threads(n: int):
    """Thread budget of the OpenMP setup helpers (SpGEMM here, BSR conversion in libfdal)."""
    lib = _host()
    lib.fdal_host_set_num_threads.argtypes = [C.c_int]
    lib.fdal_host_set_num_threads.restype = None
    lib.fdal_host_set_num_threads(int(n))


def _log(msg):
    if _VERBOSE:
        sys.stderr.write(f"[setup] {msg}\n")
        sys.stderr.flush()


def _spgemm(A: sp.csr_matrix, B: sp.csr_matrix) -> sp.csr_matrix:
    """C = A @ B (setup phase).  Large products use the OpenMP Gustavson kernel in
    csrc/host_setup.c (scipy's csr_matmat is single-threaded: minutes on the 10^8..10^9
    non-zero fine levels); small ones scipy.  cuSPARSE SpGEMM through torch was tried and
    dropped: it runs out of workspace ("insufficient resources") beyond ~3e7 non-zeros."""
    t0 = time.perf_counter()
    if A.nnz >= 2_000_000:
        A = A.tocsr()
        B = B.tocsr()
        p64, p32, pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
        Ap = np.ascontiguousarray(A.indptr, dtype=np.int64)
        Aj = np.ascontiguousarray(A.indices, dtype=np.int32)
        Ax = np.ascontiguousarray(A.data, dtype=np.float64)
        Bp = np.ascontiguousarray(B.indptr, dtype=np.int64)
        Bj = np.ascontiguousarray(B.indices, dtype=np.int32)
        Bx = np.ascontiguousarray(B.data, dtype=np.float64)
        Cp = np.empty(A.shape[0] + 1, dtype=np.int64)
        lib = _host()
        nnz = lib.fdal_host_spgemm_symbolic(A.shape[0], B.shape[1], Ap.ctypes.data_as(p64), Aj.ctypes.data_as(p32),
                                            Bp.ctypes.data_as(p64), Bj.ctypes.data_as(p32), Cp.ctypes.data_as(p64))
        Cj = np.empty(max(nnz, 1), dtype=np.int32)
        Cx = np.empty(max(nnz, 1), dtype=np.float64)
        lib.fdal_host_spgemm_numeric(A.shape[0], B.shape[1], Ap.ctypes.data_as(p64), Aj.ctypes.data_as(p32),
                                     Ax.ctypes.data_as(pd), Bp.ctypes.data_as(p64), Bj.ctypes.data_as(p32),
                                     Bx.ctypes.data_as(pd), Cp.ctypes.data_as(p64), Cj.ctypes.data_as(p32),
                                     Cx.ctypes.data_as(pd))
        C_ = sp.csr_matrix((Cx[:nnz], Cj[:nnz], Cp), shape=(A.shape[0], B.shape[1]))
        C_.has_sorted_indices = True
        _log(f"spgemm(omp) {A.shape}x{B.shape} nnz {A.nnz}x{B.nnz}->{C_.nnz}: {time.perf_counter()-t0:.2f}s")
        return C_
    C_ = (A @ B).tocsr()
    _log(f"spgemm(cpu) {A.shape}x{B.shape} nnz {A.nnz}x{B.nnz}->{C_.nnz}: {time.perf_counter()-t0:.2f}s")
    return C_


@dataclass
class Level:
    A: sp.csr_matrix
    P: sp.csr_matrix | None = None  # n_l x n_{l+1}
    R: sp.csr_matrix | None = None  # n_{l+1} x n_l
    inv_diag: np.ndarray | None = None
    lambda_max: float = 1.0


@dataclass
class Hierarchy:
    levels: list = field(default_factory=list)
    cheb_degree: int = 2  # smoother_sweeps = 2 (utilities.h:311)
    eig_ratio: float = 10.0  # "smoother: Chebyshev alpha" = 10

    def operator_complexity(self):
        return sum(L.A.nnz for L in self.levels) / self.levels[0].A.nnz

    def describe(self):
        return [(L.A.shape[0], L.A.nnz) for L in self.levels]


def _strength_graph(A: sp.csr_matrix, theta: float, comp: np.ndarray | None):
    """Symmetric SA strength: |a_ij| >= theta * sqrt(|a_ii a_jj|), i != j, same component."""
    A = A.tocoo()
    d = np.abs(sp.csr_matrix(A).diagonal())
    keep = A.row != A.col
    keep &= np.abs(A.data) >= theta * np.sqrt(d[A.row] * d[A.col])
    keep &= A.data != 0.0
    if comp is not None:
        keep &= comp[A.row] == comp[A.col]
    S = sp.csr_matrix((np.ones(keep.sum(), dtype=np.int8), (A.row[keep], A.col[keep])), shape=A.shape)
    S = ((S + S.T) > 0).astype(np.int8).tocsr()
    S.sort_indices()
    return S


def _cuda_ok(nnz: int) -> bool:
    if nnz < _GPU_SPGEMM_MIN_NNZ:
        return False
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def _to_torch_csr(M: sp.csr_matrix):
    import torch

    M = M.tocsr()
    return torch.sparse_csr_tensor(
        torch.from_numpy(M.indptr.astype(np.int64)).cuda(), torch.from_numpy(M.indices.astype(np.int64)).cuda(),
        torch.from_numpy(M.data.astype(np.float64)).cuda(), size=M.shape)


def _strength_graph_gpu(A: sp.csr_matrix, theta: float, comp: np.ndarray | None):
    """Same criterion as _strength_graph, evaluated on the GPU for the 10^8..10^9 non-zero
    fine levels (A is symmetric, so the selected pattern already is)."""
    import torch

    n = A.shape[0]
    crow = torch.from_numpy(A.indptr.astype(np.int64)).cuda()
    col = torch.from_numpy(A.indices.astype(np.int64)).cuda()
    val = torch.from_numpy(A.data).cuda()
    row = torch.repeat_interleave(torch.arange(n, device="cuda"), crow[1:] - crow[:-1])
    d = torch.from_numpy(np.abs(A.diagonal())).cuda()
    keep = (row != col) & (val.abs() >= theta * torch.sqrt(d[row] * d[col])) & (val != 0)
    if comp is not None:
        cc = torch.from_numpy(comp.astype(np.int64)).cuda()
        keep &= cc[row] == cc[col]
    rk, ck = row[keep], col[keep]
    cnt = torch.bincount(rk, minlength=n)
    indptr = np.zeros(n + 1, dtype=np.int64)
    indptr[1:] = torch.cumsum(cnt, 0).cpu().numpy()
    S = sp.csr_matrix((np.ones(int(indptr[-1]), dtype=np.int8), ck.to(torch.int32).cpu().numpy(), indptr), shape=(n, n))
    del row, col, val, keep, rk, ck
    torch.cuda.empty_cache()
    return S


def aggregate(S: sp.csr_matrix) -> tuple[np.ndarray, int]:
    n = S.shape[0]
    indptr = np.ascontiguousarray(S.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(S.indices, dtype=np.int32)
    agg = np.empty(n, dtype=np.int32)
    n_agg = _host().fdal_host_aggregate(
        n,
        indptr.ctypes.data_as(C.POINTER(C.c_int64)),
        indices.ctypes.data_as(C.POINTER(C.c_int32)),
        agg.ctypes.data_as(C.POINTER(C.c_int32)),
    )
    return agg, int(n_agg)


def estimate_lambda_max(A: sp.csr_matrix, inv_diag: np.ndarray, iters: int = 20) -> float:
    """lambda_max(D^-1 A) via a few Lanczos steps on D^-1/2 A D^-1/2 (deterministic start)."""
    n = A.shape[0]
    s = np.sqrt(np.abs(inv_diag))
    if n <= 3:
        M = (sp.diags(s) @ A @ sp.diags(s)).toarray()
        return float(np.max(np.linalg.eigvalsh(0.5 * (M + M.T))))
    if _cuda_ok(A.nnz) and not os.environ.get("FDAL_DETERMINISTIC_SETUP"):
        import torch

        At = _to_torch_csr(A)
        st = torch.from_numpy(s).cuda()

        def mv(x):
            xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).ravel()).cuda()
            return (st * (At @ (st * xt))).cpu().numpy()
    else:
        def mv(x):
            return s * (A @ (s * x))
    op = spla.LinearOperator((n, n), matvec=mv, dtype=np.float64)
    v0 = 1.0 + 0.5 * np.sin(np.arange(n) * 0.7853981633974483 + 0.3)
    try:
        lam = spla.eigsh(op, k=1, which="LA", v0=v0, ncv=min(n - 1, max(iters, 4)), maxiter=50, tol=1e-3,
                         return_eigenvectors=False)
        return float(lam[0])
    except spla.ArpackNoConvergence as e:  # best available Ritz value
        if len(e.eigenvalues):
            return float(np.max(e.eigenvalues))
        # power iteration fallback
        x = v0 / np.linalg.norm(v0)
        lam = 1.0
        for _ in range(30):
            y = op.matvec(x)
            lam = float(np.linalg.norm(y))
            x = y / lam
        return lam


def build_hierarchy(
    A: sp.csr_matrix,
    theta: float = 1e-4,
    comp: np.ndarray | None = None,
    max_coarse: int = 2000,
    max_levels: int = 12,
    cheb_degree: int = 2,
    eig_ratio: float = 10.0,
    omega: float = 4.0 / 3.0,
    verbose: bool = False,
) -> Hierarchy:
    """Smoothed-aggregation hierarchy for the explicit augmented matrix ``A``.

    ``theta``: aggregation_threshold (1e-4 default, 0.02 Stokes utilities.h:314,
    1e-3 elliptic utilities.h:731); ``comp``: per-DoF component id standing in for
    ``constant_modes`` (utilities.h:304-309) — aggregates never mix components, so
    the tentative prolongator spans the component-wise constants.
    """
    A = sp.csr_matrix(A)
    A.sort_indices()
    H = Hierarchy(cheb_degree=cheb_degree, eig_ratio=eig_ratio)
    while True:
        n = A.shape[0]
        t0 = time.perf_counter()
        diag = A.diagonal()
        inv_diag = 1.0 / diag
        lam = estimate_lambda_max(A, inv_diag)
        _log(f"level {len(H.levels)}: n={n} nnz={A.nnz} lambda_max {time.perf_counter()-t0:.2f}s")
        L = Level(A=A, inv_diag=inv_diag, lambda_max=lam)
        H.levels.append(L)
        if n <= max_coarse or len(H.levels) >= max_levels:
            break
        t0 = time.perf_counter()
        S = _strength_graph_gpu(A, theta, comp) if _cuda_ok(A.nnz) else _strength_graph(A, theta, comp)
        t1 = time.perf_counter()
        agg, n_agg = aggregate(S)
        _log(f"  strength {t1-t0:.2f}s aggregate {time.perf_counter()-t1:.2f}s -> {n_agg}")
        if n_agg >= n or n_agg == 0:  # no coarsening possible
            break
        live = np.nonzero(agg >= 0)[0]
        cnt = np.bincount(agg[live], minlength=n_agg).astype(np.float64)
        T = sp.csr_matrix((1.0 / np.sqrt(cnt[agg[live]]), (live, agg[live])), shape=(n, n_agg))
        AT = _spgemm(A, T)
        t0 = time.perf_counter()
        P = (T - (omega / lam) * (sp.diags(inv_diag) @ AT)).tocsr()
        del AT
        P.sort_indices()
        R = P.T.tocsr()
        R.sort_indices()
        _log(f"  P, R: {time.perf_counter()-t0:.2f}s")
        Ac = _spgemm(R, _spgemm(A, P))
        Ac.sort_indices()
        L.P, L.R = P, R
        if comp is not None:
            # component of an aggregate = component of its members
            cc = np.zeros(n_agg, dtype=comp.dtype)
            cc[agg[live]] = comp[live]
            comp = cc
        if verbose:
            print(f"  AMG level {len(H.levels)-1}: n={n} nnz={A.nnz} -> n_c={n_agg} lambda_max={lam:.4f}")
        A = Ac
    return H
