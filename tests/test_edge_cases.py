"""Edge cases of the boundary: malformed / ragged CSR input, call-order errors, empty rows.
The same checks run against the oracle (CPU suite) and the CUDA library (GPU suite)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import ALConfig, ALContext, FdalError
from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from . import problems as P


def _make(kind, oracle_mod, cfg):
    return oracle_mod.OracleContext(cfg) if kind == "oracle" else ALContext(cfg)


KINDS = ["oracle", pytest.param("cuda", marks=[pytest.mark.gpu])]


def _raw_set_csr(ctx, mid, nr, nc, rp, ci, v):
    rp = np.asarray(rp, dtype=np.int64)
    ci = np.asarray(ci, dtype=np.int32)
    v = np.asarray(v, dtype=np.float64)
    return ctx.api.set_csr(ctx._h, mid, nr, nc, v.size, rp.ctypes.data_as(C.POINTER(C.c_int64)),
                           ci.ctypes.data_as(C.POINTER(C.c_int32)), b.dptr(v))


@pytest.mark.parametrize("kind", KINDS)
def test_malformed_csr_is_rejected_with_a_status_code(kind, oracle_mod):
    ctx = _make(kind, oracle_mod, ALConfig())
    assert _raw_set_csr(ctx, b.MAT_A, 2, 2, [0, 1, 2], [0, 5], [1.0, 1.0]) == b.ERR_SHAPE  # column out of range
    assert _raw_set_csr(ctx, b.MAT_A, 2, 2, [0, 1, 3], [0, 1], [1.0, 1.0]) == b.ERR_SHAPE  # row_ptr[n] != nnz
    assert _raw_set_csr(ctx, b.MAT_A, 2, 2, [1, 1, 2], [0, 1], [1.0, 1.0]) == b.ERR_SHAPE  # row_ptr[0] != 0
    assert _raw_set_csr(ctx, 99, 2, 2, [0, 1, 2], [0, 1], [1.0, 1.0]) == b.ERR_INVALID  # unknown matrix id
    assert ctx.api.last_error(ctx._h)  # a message is available
    assert _raw_set_csr(ctx, b.MAT_A, 2, 2, [0, 1, 2], [0, 1], [1.0, 1.0]) == b.OK


@pytest.mark.parametrize("kind", KINDS)
def test_call_order_errors(kind, oracle_mod):
    prob, H = P.get("laplace_diag")
    ctx = _make(kind, oracle_mod, prob.config)
    x = np.zeros(prob.n_dofs)
    with pytest.raises(FdalError) as e:  # apply before finalize
        ctx.N = prob.n_dofs
        ctx.apply_system(x)
    assert e.value.status == b.ERR_STATE
    ctx.set_csr(b.MAT_A, prob.A)
    with pytest.raises(FdalError) as e:  # Ct missing
        ctx.finalize()
    assert e.value.status == b.ERR_STATE
    ctx2 = _make(kind, oracle_mod, prob.config)
    ctx2.set_csr(b.MAT_A, prob.A)
    ctx2.set_csr(b.MAT_CT, prob.Ct)
    ctx2.set_csr(b.MAT_M, prob.M)
    with pytest.raises(FdalError) as e:  # diagonal W^-1 configured but never set
        ctx2.finalize()
    assert e.value.status == b.ERR_STATE
    ctx3 = _make(kind, oracle_mod, prob.config)
    ctx3.set_csr(b.MAT_A, prob.A)
    ctx3.set_csr(b.MAT_CT, prob.Ct[:-3])  # wrong number of rows
    ctx3.set_diag(b.DIAG_W_INV, prob.winv_diag)
    with pytest.raises(FdalError) as e:
        ctx3.finalize()
    assert e.value.status == b.ERR_SHAPE


@pytest.mark.parametrize("kind", KINDS)
def test_empty_and_ragged_rows(kind, oracle_mod):
    """Rows without entries (most rows of Ct), a single dense-ish row, unsorted columns with the
    diagonal first (deal.II's layout): vmult / Tvmult equal scipy."""
    prob, H = P.get("laplace_diag")
    rng = np.random.default_rng(7)
    n, m = prob.Ct.shape
    Ct = prob.Ct.tolil()
    Ct[5, :] = rng.uniform(-1, 1, m)  # one long row among thousands of empty ones
    Ct = sp.csr_matrix(Ct)
    # deal.II ordering of A: diagonal first, then ascending columns
    A = prob.A.tocsr()
    rp, ci, v = A.indptr.copy(), A.indices.copy(), A.data.copy()
    for i in range(n):
        s, e = rp[i], rp[i + 1]
        k = s + int(np.nonzero(ci[s:e] == i)[0][0])
        ci[s:k + 1] = np.roll(ci[s:k + 1], 1)
        v[s:k + 1] = np.roll(v[s:k + 1], 1)
    A_dealii = sp.csr_matrix((v, ci, rp), shape=A.shape)
    assert not A_dealii.has_sorted_indices or True
    ctx = _make(kind, oracle_mod, prob.config)
    ctx.set_csr(b.MAT_A, A_dealii)
    ctx.set_csr(b.MAT_CT, Ct)
    ctx.set_csr(b.MAT_M, prob.M)
    ctx.set_diag(b.DIAG_W_INV, prob.winv_diag)
    ctx.set_amg(b.AMG_A11, H[b.AMG_A11])
    ctx.finalize()
    x = P.rand(n, 1)
    lam = P.rand(m, 2)
    assert P.relerr(ctx.spmv(b.MAT_A, x, n_out=n), prob.A @ x) < 1e-14
    assert P.relerr(ctx.spmv(b.MAT_CT, lam, n_out=n), Ct @ lam) < 1e-14
    assert P.relerr(ctx.spmv(b.MAT_CT, x, transpose=True, n_out=m), Ct.T @ x) < 1e-13
    g = prob.config.gamma
    ref = prob.A @ x + g * (Ct @ (prob.winv_diag * (Ct.T @ x)))
    assert P.relerr(ctx.apply_aug(x), ref) < 1e-12


# ---- CUDA library only: state and hierarchy validation -------------------------------
@pytest.mark.gpu
def test_setters_after_finalize_are_refused():
    """A finalized context owns device copies and captured graphs: every setter now returns
    FDAL_ERR_STATE (instead of silently un-finalizing a context that can never be finalized again);
    the solve entry points keep working."""
    prob, H = P.get("laplace_diag")
    ctx = syn.setup_context(ALContext(prob.config), prob, H)
    x0, _ = ctx.solve(P.rhs_of(ctx, prob))
    for call in (lambda: ctx.set_csr(b.MAT_CT, prob.Ct), lambda: ctx.set_diag(b.DIAG_W_INV, prob.winv_diag),
                 lambda: ctx.set_amg(b.AMG_A11, H[b.AMG_A11])):
        with pytest.raises(FdalError) as e:
            call()
        assert e.value.status == b.ERR_STATE and "finalized" in str(e.value)
    x1, info = ctx.solve(P.rhs_of(ctx, prob))
    assert info.status == 0 and np.array_equal(x0, x1)


def _view(A):
    return b.csr_view(sp.csr_matrix(A))


@pytest.mark.gpu
def test_malformed_hierarchy_is_rejected():
    """fdal_amg_set_level validates its CSR views like fdal_set_csr; fdal_finalize checks the shapes
    between levels and the Chebyshev parameters before anything is launched."""
    prob, H = P.get("laplace_diag")
    L0, L1 = H[b.AMG_A11].levels[0], H[b.AMG_A11].levels[1]

    def fresh():
        c = ALContext(prob.config)
        c.set_csr(b.MAT_A, prob.A)
        c.set_csr(b.MAT_CT, prob.Ct)
        c.set_csr(b.MAT_M, prob.M)
        c.set_diag(b.DIAG_W_INV, prob.winv_diag)
        return c

    def set_level(c, level, A, Pm, R, degree=2, lmax=1.5, ratio=10.0):
        Av, ka = _view(A)
        Pv, kp = _view(Pm) if Pm is not None else (None, None)
        Rv, kr = _view(R) if R is not None else (None, None)
        return c.api.amg_set_level(c._h, b.AMG_A11, level, C.byref(Av), C.byref(Pv) if Pv is not None else None,
                                   C.byref(Rv) if Rv is not None else None, None, lmax, degree, ratio)

    # a view with a column index out of range never reaches the device
    c = fresh()
    bad = sp.csr_matrix(L0.P).copy()
    bad.indices = bad.indices.copy()
    bad.indices[0] = bad.shape[1] + 7
    Av, ka = _view(L0.A)
    rp, ci, v = b.csr_arrays(bad)
    Pv = b.CsrView(bad.shape[0], bad.shape[1], v.size, rp.ctypes.data_as(C.POINTER(C.c_int64)),
                   ci.ctypes.data_as(C.POINTER(C.c_int32)), v.ctypes.data_as(C.POINTER(C.c_double)))
    assert c.api.amg_set_level(c._h, b.AMG_A11, 0, C.byref(Av), C.byref(Pv), None, None, 1.5, 2, 10.0) == b.ERR_SHAPE
    # Chebyshev degree 0 on a smoothed level (the V-cycle would never write its result)
    c = fresh()
    assert set_level(c, 0, L0.A, L0.P, L0.R, degree=0) == b.OK
    assert set_level(c, 1, L1.A, None, None) == b.OK
    with pytest.raises(FdalError) as e:
        c.finalize()
    assert e.value.status == b.ERR_INVALID
    # R with the wrong shape
    c = fresh()
    assert set_level(c, 0, L0.A, L0.P, sp.csr_matrix(L0.R)[:-1]) == b.OK
    assert set_level(c, 1, L1.A, None, None) == b.OK
    with pytest.raises(FdalError) as e:
        c.finalize()
    assert e.value.status == b.ERR_SHAPE
    # P whose column count is not the next level's row count
    c = fresh()
    assert set_level(c, 0, L0.A, sp.csr_matrix(L0.P)[:, :-1], None) == b.OK
    assert set_level(c, 1, L1.A, None, None) == b.OK
    with pytest.raises(FdalError) as e:
        c.finalize()
    assert e.value.status == b.ERR_SHAPE
