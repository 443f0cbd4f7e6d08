"""Merge-path CSR SpMV (k_spmv_merge + carry fix-up, north_star "warp-per-row and merge-path variants"):
forced on for every eligible matrix (FDAL_MERGE=1) it must reproduce scipy on awkward shapes — rows
longer than a CTA tile, thousands of empty rows, a single row, fewer rows than threads — and the whole
solve (restriction and prolongation then run through it) must still match the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import ALContext
from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from . import parity_log as PL
from . import problems as P

pytestmark = pytest.mark.gpu


def _ctx_with(prob, H, Ct, monkeypatch, merge="1"):
    monkeypatch.setenv("FDAL_MERGE", merge)
    ctx = ALContext(prob.config)
    ctx.set_csr(b.MAT_A, prob.A)
    ctx.set_csr(b.MAT_CT, Ct)
    ctx.set_csr(b.MAT_M, prob.M)
    ctx.set_diag(b.DIAG_W_INV, prob.winv_diag)
    ctx.set_amg(b.AMG_A11, H[b.AMG_A11])
    ctx.finalize()
    return ctx


@pytest.mark.parametrize("shape", ["as_is", "long_rows", "one_dense_column", "mostly_empty", "dense_block"])
def test_merge_path_spmv_matches_scipy(shape, monkeypatch):
    prob, H = P.get("laplace_diag")
    rng = np.random.default_rng(11)
    n, m = prob.Ct.shape
    Ct = prob.Ct.tolil()
    if shape == "long_rows":  # rows of C (= columns of Ct) far longer than one CTA tile of 2048 merge items
        for j in (0, m // 2, m - 1):
            Ct[:, j] = rng.uniform(-1, 1, (n, 1))
    elif shape == "one_dense_column":
        Ct = sp.lil_matrix((n, m))
        Ct[:, 3] = rng.uniform(-1, 1, (n, 1))
    elif shape == "mostly_empty":
        Ct = sp.lil_matrix((n, m))
        Ct[7, 1] = 2.0
        Ct[n - 1, m - 1] = -3.0
    elif shape == "dense_block":
        Ct[100:140, :] = rng.uniform(-1, 1, (40, m))
    Ct = sp.csr_matrix(Ct)
    ctx = _ctx_with(prob, H, Ct, monkeypatch)
    x, lam = P.rand(n, 1), P.rand(m, 2)
    PL.check(f"merge-path A x [{shape}]", P.relerr(ctx.spmv(b.MAT_A, x, n_out=n), prob.A @ x), 1e-14)
    PL.check(f"merge-path Ct lam [{shape}]", P.relerr(ctx.spmv(b.MAT_CT, lam, n_out=n), Ct @ lam), 1e-14)
    ref = Ct.T @ x
    got = ctx.spmv(b.MAT_CT, x, transpose=True, n_out=m)
    PL.check(f"merge-path C x [{shape}]", float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)), 1e-13)
    g = prob.config.gamma
    ref = prob.A @ x + g * (Ct @ (prob.winv_diag * (Ct.T @ x)))
    PL.check(f"apply_aug with merge-path C x [{shape}]", P.relerr(ctx.apply_aug(x), ref), 1e-12)
    ctx.close()


@pytest.mark.parametrize("name", ["laplace_diag", "stokes2d_diag", "elliptic_modified_diag", "stokes3d_diag"])
def test_solve_with_merge_path_everywhere(name, oracle_mod, monkeypatch):
    monkeypatch.setenv("FDAL_MERGE", "1")
    prob, H = P.get(name)
    gpu = syn.setup_context(ALContext(prob.config), prob, H)
    ora = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    r = P.rand(prob.sizes[0], 7)
    PL.check("apply_amg (R, P through merge-path)", P.relerr(gpu.apply_amg(r), ora.apply_amg(r)), 1e-12)
    X = P.rand(prob.n_dofs, 5)
    PL.check("apply_system", P.relerr(gpu.apply_system(X), ora.apply_system(X)), 1e-12)
    rhs = P.rhs_of(ora, prob)
    xg, ig = gpu.solve(rhs)
    xo, io = ora.solve(rhs)
    assert abs(ig.outer_iterations - io.outer_iterations) <= 1
    if ig.outer_iterations == io.outer_iterations:
        PL.check("solve: solution", P.relerr(xg, xo), 1e-8)
