"""CSR -> BSR conversion on the device (csrc/bsr_build.cu, the routine fdal_finalize runs for the dim-blocked
matrices) against the host conversion (csrc/host_finalize.h) bit for bit, and against scipy."""
import copy
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import ALContext, amg_setup
from fictitious_domain_al_preconditioners_b200 import partition as part
from fictitious_domain_al_preconditioners_b200.context import csr_to_bsr

from . import problems as P
from .test_host_finalize import _deal_ii_order, _p, _ptrs

pytestmark = pytest.mark.gpu


def _host_bsr(A, b, max_fill=1.35):
    L = C.CDLL(amg_setup.build_host_lib())
    p64, p32, pd = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    L.fdal_hostfin_bsr.restype = C.c_int64
    L.fdal_hostfin_bsr.argtypes = [C.c_int64, C.c_int64, C.c_int64, p64, p32, pd, C.c_int32, C.c_double, p32, p32, pd]
    rp, ci, v = _ptrs(A)
    n = A.shape[0]
    brp = np.empty(n // b + 1, dtype=np.int32)
    args = (n, A.shape[1], A.nnz, _p(rp, C.c_int64), _p(ci, C.c_int32), _p(v, C.c_double), b, max_fill)
    nblk = L.fdal_hostfin_bsr(*args, _p(brp, C.c_int32), None, None)
    if nblk < 0:
        return None
    bcj = np.empty(nblk, dtype=np.int32)
    bv = np.empty(nblk * b * b, dtype=np.float64)
    L.fdal_hostfin_bsr(*args, _p(brp, C.c_int32), _p(bcj, C.c_int32), _p(bv, C.c_double))
    return brp, bcj, bv.reshape(nblk, b, b)


def _blocked(nb, b, density, seed, knock_out=True):
    rng = np.random.default_rng(seed)
    pattern = sp.random(nb, nb, density=density, random_state=seed, format="csr") + sp.identity(nb, format="csr")
    A = sp.kron(pattern, np.ones((b, b)), format="csr")
    A.data = rng.uniform(-1, 1, A.nnz)
    if knock_out:  # partly filled blocks
        A.data[rng.choice(A.nnz, A.nnz // 7, replace=False)] = 0.0
        A.eliminate_zeros()
    return _deal_ii_order(A)


@pytest.mark.parametrize("b,nb,density", [(2, 500, 0.02), (3, 400, 0.03), (3, 64, 0.9), (2, 3000, 0.001)])
def test_device_conversion_equals_host_conversion(b, nb, density):
    A = _blocked(nb, b, density, seed=b * 100 + nb)
    dev = csr_to_bsr(A, b)
    host = _host_bsr(A, b)
    assert dev is not None and host is not None
    for d, h in zip(dev, host):
        assert np.array_equal(d, h)
    ref = sp.csr_matrix(A)
    ref.sort_indices()
    ref = ref.tobsr(blocksize=(b, b))
    ref.sort_indices()
    assert np.array_equal(dev[0], ref.indptr) and np.array_equal(dev[1], ref.indices) and np.array_equal(dev[2], ref.data)


def test_long_block_rows_and_empty_block_rows():
    """One block row with 7 800 scalar entries (the hash set is sized by the longest block row: 16 384 slots, two
    block rows in flight per CTA instead of four), many empty ones."""
    b, nb = 3, 1500
    rng = np.random.default_rng(1)
    A = sp.lil_matrix((nb * b, nb * b))
    cols = rng.choice(nb * b, 2600, replace=False)
    for r in range(b):
        A[7 * b + r, cols] = rng.uniform(-1, 1, cols.size)
    for I in range(0, nb, 3):
        A[I * b:(I + 1) * b, I * b:(I + 1) * b] = rng.uniform(1, 2, (b, b))
    A = sp.csr_matrix(A)
    dev = csr_to_bsr(A, b, max_fill=10.0)
    host = _host_bsr(A, b, max_fill=10.0)
    for d, h in zip(dev, host):
        assert np.array_equal(d, h)
    assert np.all(np.diff(dev[0])[[1, 2, 4, 5, 8]] == 0)  # empty block rows (every third block row has a diagonal block)


def test_block_rows_beyond_the_shared_memory_set_are_left_to_the_host():
    """27 000 scalar entries in one block row would need a 65 536-slot set (393 KB for one warp): the device routine
    declines, the host conversion (the fallback inside fdal_finalize) still blocks the matrix."""
    b, nb = 3, 4000
    rng = np.random.default_rng(2)
    A = sp.lil_matrix((nb * b, nb * b))
    cols = rng.choice(nb * b, 9000, replace=False)
    for r in range(b):
        A[5 * b + r, cols] = 1.0
    A = sp.csr_matrix(A + sp.identity(nb * b))
    assert csr_to_bsr(A, b, max_fill=100.0) is None
    assert _host_bsr(A, b, max_fill=100.0) is not None


def test_device_conversion_declines_like_the_host():
    A = sp.identity(300, format="csr")  # scalar diagonal: 3x3 blocking stores 3x the non-zeros
    assert csr_to_bsr(A, 3) is None and _host_bsr(A, 3) is None
    assert csr_to_bsr(A, 3, max_fill=4.0) is not None


@pytest.mark.parametrize("name", ["stokes2d_node", "stokes3d_node", "elasticity"])
def test_finalize_converts_on_the_device(name, monkeypatch):
    """fdal_finalize converts the velocity / elasticity block and the finest AMG operator of a node-numbered
    problem on the device; FDAL_HOST_BSR=1 (the OpenMP conversion) gives bit-identical operators."""
    prob, H = P.get(name)
    lp = part.distribute_problem(prob, H, 0, 1)
    cfg = copy.deepcopy(prob.config)
    cfg.block_size = lp.block_size
    assert cfg.block_size in (2, 3)
    gpu = part.setup_local_context(ALContext(cfg), lp)
    on_dev, on_host = gpu.bsr_conversions()
    assert on_dev >= 1 and on_host == 0, (on_dev, on_host)
    X = lp.scatter(P.rand(prob.n_dofs, 5))
    r = X[: prob.sizes[0]].copy()
    y_dev, z_dev = gpu.apply_system(X), gpu.apply_amg(r)
    monkeypatch.setenv("FDAL_HOST_BSR", "1")
    ref = part.setup_local_context(ALContext(cfg), lp)
    assert ref.bsr_conversions() == (0, on_dev)
    assert np.array_equal(y_dev, ref.apply_system(X))
    assert np.array_equal(z_dev, ref.apply_amg(r))
