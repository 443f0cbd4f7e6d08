"""The host half of fdal_finalize (csrc/host_finalize.h, compiled into libfdal_host.so through
csrc/host_finalize_hooks.cpp): the stable transpose that turns SparseMatrix::Tvmult into a gather
(C = Ct^T, B = Bt^T, R = P^T) — serial and OpenMP form — and the CSR -> BSR conversion, against scipy."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import amg_setup


@pytest.fixture(scope="module")
def lib():
    L = C.CDLL(amg_setup.build_host_lib())
    p64, p32, pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.fdal_hostfin_transpose.restype = None
    L.fdal_hostfin_transpose.argtypes = [C.c_int64, C.c_int64, C.c_int64, p64, p32, pd, C.c_int64, p32, p32, pd]
    L.fdal_hostfin_bsr.restype = C.c_int64
    L.fdal_hostfin_bsr.argtypes = [C.c_int64, C.c_int64, C.c_int64, p64, p32, pd, C.c_int32, C.c_double, p32, p32, pd]
    return L


def _ptrs(A):
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    v = np.ascontiguousarray(A.data, dtype=np.float64)
    return rp, ci, v


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _transpose(lib, A, min_parallel_nnz):
    rp, ci, v = _ptrs(A)
    nr, nc = A.shape
    t_rp = np.empty(nc + 1, dtype=np.int32)
    t_ci = np.empty(max(A.nnz, 1), dtype=np.int32)
    t_v = np.empty(max(A.nnz, 1), dtype=np.float64)
    lib.fdal_hostfin_transpose(nr, nc, A.nnz, _p(rp, C.c_int64), _p(ci, C.c_int32), _p(v, C.c_double), min_parallel_nnz,
                               _p(t_rp, C.c_int32), _p(t_ci, C.c_int32), _p(t_v, C.c_double))
    return t_rp, t_ci[: A.nnz], t_v[: A.nnz]


def _deal_ii_order(A):
    """diagonal first, rest ascending (deal.II's row layout): rows are NOT sorted"""
    A = sp.csr_matrix(A)
    A.sort_indices()
    rp, ci, v = A.indptr, A.indices.copy(), A.data.copy()
    for i in range(min(A.shape)):
        s, e = rp[i], rp[i + 1]
        hit = np.nonzero(ci[s:e] == i)[0]
        if hit.size:
            k = s + int(hit[0])
            ci[s:k + 1] = np.roll(ci[s:k + 1], 1)
            v[s:k + 1] = np.roll(v[s:k + 1], 1)
    return sp.csr_matrix((v, ci, rp), shape=A.shape)


@pytest.mark.parametrize("min_parallel_nnz", [1 << 40, 0], ids=["serial", "openmp"])
@pytest.mark.parametrize("shape,density", [((700, 300), 0.02), ((50, 4000), 0.01), ((2000, 2000), 0.004), ((40, 40), 0.0)])
def test_transpose_is_the_stable_counting_sort(lib, shape, density, min_parallel_nnz):
    A = sp.random(*shape, density=density, random_state=3, format="csr")
    if shape[0] == shape[1] and density > 0:
        A = _deal_ii_order(A + sp.identity(shape[0], format="csr"))
    t_rp, t_ci, t_v = _transpose(lib, A, min_parallel_nnz)
    # reference: entries of column j in the order rows are walked (stable), whatever the order inside A's rows
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    order = np.argsort(A.indices, kind="stable")
    assert np.array_equal(t_rp, np.concatenate([[0], np.cumsum(np.bincount(A.indices, minlength=A.shape[1]))]))
    assert np.array_equal(t_ci, rows[order])
    assert np.array_equal(t_v, A.data[order])
    T = sp.csr_matrix((t_v, t_ci, t_rp), shape=A.shape[::-1])
    assert abs(T - A.T).max() == 0 if A.nnz else T.nnz == 0


def test_transpose_with_empty_rows_and_uneven_columns(lib):
    """Ct of an immersed-boundary problem: almost every row empty, a few columns hold everything (the task
    cuts of the OpenMP form land inside long runs of equal row pointers)."""
    rng = np.random.default_rng(0)
    n, m = 5000, 37
    rows = rng.choice(n, 150, replace=False)
    A = sp.lil_matrix((n, m))
    for r in rows:
        cols = rng.choice(m, 5, replace=False)
        A[r, cols] = rng.uniform(-1, 1, 5)
    A[rows[:120], 3] = 1.5  # one heavy column
    A = sp.csr_matrix(A)
    for mp in (1 << 40, 0):
        t_rp, t_ci, t_v = _transpose(lib, A, mp)
        T = sp.csr_matrix((t_v, t_ci, t_rp), shape=(m, n))
        assert abs(T - A.T).max() == 0
        assert all(np.all(np.diff(t_ci[t_rp[j]:t_rp[j + 1]]) > 0) for j in range(m))


@pytest.mark.parametrize("b", [2, 3])
def test_bsr_conversion_matches_scipy(lib, b):
    rng = np.random.default_rng(b)
    nb = 240
    pattern = sp.random(nb, nb, density=0.03, random_state=5, format="csr") + sp.identity(nb, format="csr")
    A = sp.kron(pattern, np.ones((b, b)), format="csr")
    A.data = rng.uniform(-1, 1, A.nnz)
    # knock out single scalar entries so that blocks are only partly filled, and use deal.II's row order
    A.data[rng.choice(A.nnz, A.nnz // 7, replace=False)] = 0.0
    A.eliminate_zeros()
    A = _deal_ii_order(A)
    rp, ci, v = _ptrs(A)
    n = A.shape[0]
    brp = np.empty(n // b + 1, dtype=np.int32)
    args = (n, n, A.nnz, _p(rp, C.c_int64), _p(ci, C.c_int32), _p(v, C.c_double), b, 1.35)
    nblk = lib.fdal_hostfin_bsr(*args, _p(brp, C.c_int32), None, None)
    assert nblk > 0
    bcj = np.empty(nblk, dtype=np.int32)
    bv = np.empty(nblk * b * b, dtype=np.float64)
    assert lib.fdal_hostfin_bsr(*args, _p(brp, C.c_int32), _p(bcj, C.c_int32), _p(bv, C.c_double)) == nblk
    ref = sp.csr_matrix(A)
    ref.sort_indices()
    ref = ref.tobsr(blocksize=(b, b))
    ref.sort_indices()
    assert np.array_equal(brp, ref.indptr) and np.array_equal(bcj, ref.indices)
    assert np.array_equal(bv.reshape(-1, b, b), ref.data)


def test_bsr_conversion_declines(lib):
    n = 12
    A = sp.identity(n, format="csr")  # a scalar diagonal: 3x3 blocking stores 3x the non-zeros
    rp, ci, v = _ptrs(A)
    brp = np.empty(n // 3 + 1, dtype=np.int32)
    assert lib.fdal_hostfin_bsr(n, n, A.nnz, _p(rp, C.c_int64), _p(ci, C.c_int32), _p(v, C.c_double), 3, 1.35,
                                _p(brp, C.c_int32), None, None) == -1
    # rows not a multiple of the block size
    B = sp.identity(13, format="csr")
    rp, ci, v = _ptrs(B)
    assert lib.fdal_hostfin_bsr(13, 13, B.nnz, _p(rp, C.c_int64), _p(ci, C.c_int32), _p(v, C.c_double), 3, 10.0,
                                _p(brp, C.c_int32), None, None) == -1


def test_chebyshev_plan_of_the_exact_mass_solve(lib):
    """The host half of the Chebyshev mass solve (csrc/host_finalize.h: Lanczos bounds from the calibration CG,
    iteration count, coefficients) on the immersed mass matrix of elliptic_interface: the Ritz interval lies inside
    the spectrum of D^-1 M and is tight, and the fixed-count recurrence with the returned coefficients solves
    M x = b to rounding."""
    import scipy.sparse.linalg as sla

    from fictitious_domain_al_preconditioners_b200 import synthetic as syn

    M = syn.elliptic_interface(cycle=3).M.tocsr()
    n = M.shape[0]
    invd = 1.0 / M.diagonal()
    b = np.random.default_rng(0).uniform(-1, 1, n)
    # Jacobi-PCG as mass_calibrate runs it: record r.z and p.Ap
    x, r, p, rho_old = np.zeros(n), b.copy(), np.zeros(n), np.inf
    rhos, pvs, rr0 = [], [], b @ b
    while np.sqrt(r @ r) > 1e-17 * np.sqrt(rr0) and len(rhos) < 300:
        z = invd * r
        rho = r @ z
        p = z + (0.0 if np.isinf(rho_old) else rho / rho_old) * p
        v = M @ p
        pv = p @ v
        x += rho / pv * p
        r -= rho / pv * v
        rho_old = rho
        rhos.append(rho)
        pvs.append(pv)
    rho_a, pv_a = np.array(rhos), np.array(pvs)
    lib.fdal_hostfin_cheb_plan.restype = C.c_int32
    lib.fdal_hostfin_cheb_plan.argtypes = [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32,
                                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lo, hi = C.c_double(), C.c_double()
    coef = np.zeros(600)
    its = lib.fdal_hostfin_cheb_plan(len(rhos), _p(rho_a, C.c_double), _p(pv_a, C.c_double), 300, C.byref(lo), C.byref(hi),
                                     _p(coef, C.c_double))
    Dh = sp.diags(np.sqrt(invd))
    ev = np.linalg.eigvalsh((Dh @ M @ Dh).toarray())
    assert lo.value <= ev[0] * 1.001 and hi.value >= ev[-1] * 0.999  # widened Ritz interval covers the spectrum
    assert lo.value >= 0.9 * ev[0] and hi.value <= 1.1 * ev[-1]      # and is tight
    assert 30 <= its <= 80 and coef[0] == 0.0
    d = coef[1] * invd * b
    y = d.copy()
    for k in range(1, its):
        d = coef[2 * k] * d + coef[2 * k + 1] * invd * (b - M @ y)
        y = y + d
    assert np.linalg.norm(b - M @ y) <= 2e-15 * np.linalg.norm(b)
    assert np.linalg.norm(y - sla.spsolve(M.tocsc(), b)) <= 1e-14 * np.linalg.norm(y)
    # an unusable history (too short / non-positive) gives no plan
    assert lib.fdal_hostfin_cheb_plan(2, _p(rho_a, C.c_double), _p(pv_a, C.c_double), 300, C.byref(lo), C.byref(hi),
                                      _p(coef, C.c_double)) == 0
    # and so does a cap below the needed count
    assert lib.fdal_hostfin_cheb_plan(len(rhos), _p(rho_a, C.c_double), _p(pv_a, C.c_double), 10, C.byref(lo), C.byref(hi),
                                      _p(coef, C.c_double)) == 0
