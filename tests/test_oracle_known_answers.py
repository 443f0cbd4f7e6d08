"""Known-answer checks that pin the CPU oracle (SURVEY.md 8(c) invariants).

The reference ships no golden vectors, so apart from the preconditioner classes
(pinned against the compiled reference header in test_reference_pinning.py) the
oracle is pinned by algebra: each piece is compared with an independent
numpy/scipy evaluation of the formula written in the reference source.
"""
import copy

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn
from fictitious_domain_al_preconditioners_b200.context import (
    IterationNumberControl,
    NoConvergence,
    ReductionControl,
    SolverControl,
)

from . import problems as P


def ctx_for(oracle_mod, prob, H, **over):
    cfg = copy.deepcopy(prob.config)
    for k, v in over.items():
        setattr(cfg, k, v)
    p2 = copy.copy(prob)
    p2.config = cfg
    return syn.setup_context(oracle_mod.OracleContext(cfg), p2, H, oracle=True), p2


def winv_matrix(prob):
    m = prob.M.shape[0]
    mode = prob.config.winv_mode
    if mode == b.WINV_DIAG:
        return np.diag(prob.winv_diag)
    Mi = np.linalg.inv(prob.M.toarray())
    return Mi if mode == b.WINV_EXACT_M else Mi @ Mi


def aug_matrices(prob):
    """Dense A11g (and A22g, A12g, A21g) exactly as written in the reference."""
    cfg = prob.config
    A = prob.A.toarray()
    Ct = prob.Ct.toarray()
    W = winv_matrix(prob)
    A11 = A if cfg.aug_explicit else A + cfg.gamma * Ct @ W @ Ct.T
    out = dict(A11=A11, Ct=Ct, W=W)
    if prob.A2 is not None:
        M = prob.M.toarray()
        out["A22"] = prob.A2.toarray() + cfg.gamma2 * M @ W @ M
        out["A12"] = -cfg.gamma * Ct @ W @ M
        out["A21"] = -cfg.gamma2 * M @ W @ Ct.T
        out["M"] = M
    return out


def system_matrix(prob):
    k = prob.config.kind
    a = aug_matrices(prob)
    n, m = prob.Ct.shape
    if k == b.KIND_LAPLACE:
        return np.block([[a["A11"], a["Ct"]], [a["Ct"].T, np.zeros((m, m))]])
    if k in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
        Bt = prob.Bt.toarray()
        npr = Bt.shape[1]
        return np.block(
            [
                [a["A11"], Bt, a["Ct"]],
                [Bt.T, np.zeros((npr, npr)), np.zeros((npr, m))],
                [a["Ct"].T, np.zeros((m, npr)), np.zeros((m, m))],
            ]
        )
    M = a["M"]
    return np.block([[a["A11"], a["A12"], a["Ct"]], [a["A21"], a["A22"], -M], [a["Ct"].T, -M, np.zeros((m, m))]])


SMALL = {
    "laplace_diag": (syn.immersed_laplace, dict(r_bg=4, diagonal_inverse=True)),
    "laplace_exact": (syn.immersed_laplace, dict(r_bg=4, diagonal_inverse=False)),
    "laplace_opform": (syn.immersed_laplace, dict(r_bg=4, operator_form=True, diagonal_inverse=False)),
    "stokes2d_exact": (syn.stokes_immersed_boundary, dict(dim=2, nel=8)),
    "stokes2d_diag": (syn.stokes_immersed_boundary, dict(dim=2, nel=8, diagonal_mass=True)),
    "stokes3d_diag": (syn.stokes_immersed_boundary, dict(dim=3, nel=4, r_emb=1)),
    "elliptic_modified": (syn.elliptic_interface, dict(cycle=0)),
    "elliptic_modified_diag": (syn.elliptic_interface, dict(cycle=0, diagonal_inverse=True)),
    "elliptic_ideal": (syn.elliptic_interface, dict(cycle=0, modified=False, gamma_solid=10.0)),
    "elliptic_m2": (syn.elliptic_interface, dict(cycle=0, h_scaled=False, diagonal_inverse=False)),
    "elasticity": (syn.elasticity_interface, dict(cycle=0)),
    "elasticity_diag": (syn.elasticity_interface, dict(cycle=1, diagonal_inverse=True)),
    "stokes2d_node": (syn.stokes_immersed_boundary, dict(dim=2, nel=8, diagonal_mass=True, numbering="node")),
}


@pytest.fixture(scope="module", params=list(SMALL))
def small(request):
    fac, kw = SMALL[request.param]
    prob = fac(**kw)
    H = syn.build_hierarchies(prob, max_coarse=60)
    return request.param, prob, H


def test_superlu_handover_is_a_direct_solve(oracle_mod):
    prob, H = P.get("laplace_exact")
    ctx, _ = ctx_for(oracle_mod, prob, H, winv_mode=b.WINV_EXACT_M)
    t = P.rand(prob.M.shape[0], 1)
    ref = spla.splu(sp.csc_matrix(prob.M)).solve(t)
    assert P.relerr(ctx.apply_winv(t), ref) < 1e-13
    ctx2, _ = ctx_for(oracle_mod, prob, H, winv_mode=b.WINV_EXACT_M_SQUARED)
    ref2 = spla.splu(sp.csc_matrix(prob.M)).solve(ref)
    assert P.relerr(ctx2.apply_winv(t), ref2) < 1e-13


def test_vmult_and_tvmult_match_scipy(oracle_mod):
    prob, H = P.get("stokes2d_diag")
    ctx, _ = ctx_for(oracle_mod, prob, H)
    for mid, A in ((b.MAT_A, prob.A), (b.MAT_CT, prob.Ct), (b.MAT_BT, prob.Bt), (b.MAT_MP, prob.Mp)):
        x = P.rand(A.shape[1], 2)
        assert P.relerr(ctx.spmv(mid, x, n_out=A.shape[0]), A @ x) < 1e-15 * 50
        xt = P.rand(A.shape[0], 3)
        assert P.relerr(ctx.spmv(mid, xt, transpose=True, n_out=A.shape[1]), A.T @ xt) < 1e-15 * 50


def test_operators_match_the_reference_formulas(small, oracle_mod):
    name, prob, H = small
    ctx, _ = ctx_for(oracle_mod, prob, H)
    a = aug_matrices(prob)
    x = P.rand(prob.sizes[0], 4)
    assert P.relerr(ctx.apply_aug(x), a["A11"] @ x) < 1e-12
    if "A22" in a:
        x2 = P.rand(prob.sizes[1], 5)
        assert P.relerr(ctx.apply_aug(x2, b.AMG_A22), a["A22"] @ x2) < 1e-12
    AA = system_matrix(prob)
    X = P.rand(prob.n_dofs, 6)
    assert P.relerr(ctx.apply_system(X), AA @ X) < 1e-12
    if prob.augment_rhs:
        n = prob.sizes[0]
        g = prob.rhs[-prob.Ct.shape[1]:]
        ref = prob.rhs.copy()
        ref[:n] += prob.config.gamma * a["Ct"] @ a["W"] @ g
        assert P.relerr(ctx.augment_rhs(prob.rhs), ref) < 1e-12


def vcycle_ref(H, bvec, l=0):
    """Independent numpy restatement of the ML V-cycle with Chebyshev smoothing
    (SURVEY.md App. A.6)."""
    L = H.levels[l]
    A = L.A
    if l == len(H.levels) - 1:
        return np.linalg.solve(A.toarray(), bvec)
    invd = L.inv_diag
    beta, alpha = 1.1 * L.lambda_max, L.lambda_max / H.eig_ratio
    delta, theta = 0.5 * (beta - alpha), 0.5 * (beta + alpha)
    s1 = theta / delta

    def cheb(x, zero):
        rho = 1.0 / s1
        d = invd * (bvec if zero else bvec - A @ x) / theta
        x = d.copy() if zero else x + d
        for _ in range(1, H.cheb_degree):
            rho1 = 1.0 / (2 * s1 - rho)
            d = rho1 * rho * d + 2 * rho1 / delta * invd * (bvec - A @ x)
            x = x + d
            rho = rho1
        return x

    x = cheb(None, True)
    r = bvec - A @ x
    e = vcycle_ref(H, L.R @ r, l + 1)
    x = x + L.P @ e
    return cheb(x, False)


def test_vcycle_matches_independent_restatement(small, oracle_mod):
    name, prob, H = small
    ctx, _ = ctx_for(oracle_mod, prob, H)
    def tol(Hh):  # two different direct coarse solves agree to eps * cond(A_coarse)
        return max(1e-11, 100 * 2.2e-16 * np.linalg.cond(Hh.levels[-1].A.toarray()))

    r = P.rand(prob.sizes[0], 7)
    assert P.relerr(ctx.apply_amg(r), vcycle_ref(H[b.AMG_A11], r)) < tol(H[b.AMG_A11])
    if b.AMG_A22 in H:
        r2 = P.rand(prob.sizes[1], 8)
        assert P.relerr(ctx.apply_amg(r2, b.AMG_A22), vcycle_ref(H[b.AMG_A22], r2)) < tol(H[b.AMG_A22])


def test_vcycle_is_a_symmetric_linear_operator(oracle_mod):
    prob, H = P.get("laplace_diag")
    ctx, _ = ctx_for(oracle_mod, prob, H)
    n = prob.sizes[0]
    x, y = P.rand(n, 9), P.rand(n, 10)
    Bx, By = ctx.apply_amg(x), ctx.apply_amg(y)
    assert abs(y @ Bx - x @ By) < 1e-12 * abs(y @ Bx)
    assert P.relerr(ctx.apply_amg(2.0 * x - 3.0 * y), 2.0 * Bx - 3.0 * By) < 1e-13


def cg_ref(Aop, Bop, rhs, ctl):
    """numpy restatement of deal.II SolverCG + SolverControl."""
    x = np.zeros_like(rhs)
    r = rhs.copy()
    res0 = np.linalg.norm(r)

    def check(step, val):
        if ctl.type == b.CONTROL_REDUCTION and val < ctl.reduce * res0:
            return 1
        if ctl.type == b.CONTROL_ITERATION_NUMBER and step >= ctl.max_steps:
            return 1
        if val <= ctl.tol:
            return 1
        return 2 if step >= ctl.max_steps else 0

    st, it, rho_old, p = check(0, res0), 0, 0.0, None
    while st == 0:
        it += 1
        z = Bop(r)
        rho = r @ z
        p = z if it == 1 else z + (rho / rho_old) * p
        v = Aop(p)
        alpha = rho / (p @ v)
        x = x + alpha * p
        r = r - alpha * v
        rho_old = rho
        st = check(it, np.sqrt(abs(r @ r)))
    return x, it, st


@pytest.mark.parametrize("ctl", [SolverControl(100, 1e-2), SolverControl(200, 1e-9), ReductionControl(1000, 1e-30, 1e-6),
                                 IterationNumberControl(7, 1e-30)])
def test_inner_cg_follows_dealii_control_semantics(ctl, oracle_mod):
    prob, H = P.get("laplace_diag")
    ctx, _ = ctx_for(oracle_mod, prob, H, inner=ctl)
    a = aug_matrices(prob)
    rhs = P.rand(prob.sizes[0], 11)
    x, its = ctx.apply_aug_inv(rhs)
    xr, itr, st = cg_ref(lambda v: a["A11"] @ v, lambda v: vcycle_ref(H[b.AMG_A11], v), rhs, ctl)
    assert st == 1 and its == itr
    assert P.relerr(x, xr) < 1e-9
    if ctl.type == b.CONTROL_ITERATION_NUMBER:
        assert its == 7


def test_inner_failure_is_no_convergence(oracle_mod):
    prob, H = P.get("laplace_diag")
    ctx, p2 = ctx_for(oracle_mod, prob, H, inner=SolverControl(2, 1e-30))
    with pytest.raises(NoConvergence) as e:
        ctx.apply_aug_inv(P.rand(prob.sizes[0], 12))
    assert e.value.status == b.ERR_INNER_NO_CONVERGENCE
    with pytest.raises(NoConvergence):
        ctx.solve(P.rhs_of(ctx, p2))


def prec_matrix(prob):
    """Dense P of each AL preconditioner, from the block algebra in
    augmented_lagrangian_preconditioner.h."""
    cfg = prob.config
    a = aug_matrices(prob)
    n, m = prob.Ct.shape
    Wg = -np.linalg.inv(a["W"]) / cfg.gamma  # (-gamma invW)^-1
    k = cfg.kind
    if k == b.KIND_LAPLACE:
        return np.block([[a["A11"], a["Ct"]], [np.zeros((m, n)), Wg]])
    if k == b.KIND_STOKES:
        Bt = prob.Bt.toarray()
        npr = Bt.shape[1]
        Mp = -prob.Mp.toarray() / cfg.gamma_grad_div
        return np.block([[a["A11"], Bt, a["Ct"]], [np.zeros((npr, n)), Mp, np.zeros((npr, m))],
                         [np.zeros((m, n)), np.zeros((m, npr)), Wg]])
    M = a["M"]
    if k == b.KIND_ELLIPTIC_IDEAL:
        return np.block([[a["A11"], a["A12"], a["Ct"]], [a["A21"], a["A22"], -M],
                         [np.zeros((m, n)), np.zeros((m, m)), Wg]])
    # modified: d0 = A11^-1 (u + gamma Ct W M d1 - Ct d2) -> block upper triangular with -A12-like term
    return np.block([[a["A11"], -cfg.gamma * a["Ct"] @ a["W"] @ M, a["Ct"]],
                     [np.zeros((m, n)), a["A22"], -M], [np.zeros((m, n)), np.zeros((m, m)), Wg]])


def test_preconditioner_is_the_inverse_of_its_block_triangular_matrix(small, oracle_mod):
    """P . vmult(u) == u when the inner solves are run to 1e-13 (invariant (i))."""
    name, prob, H = small
    tight = SolverControl(2000, 1e-13)
    over = dict(inner=tight)
    if prob.config.kind in (b.KIND_STOKES,):
        over["mass"] = SolverControl(500, 1e-14)
    ctx, p2 = ctx_for(oracle_mod, prob, H, **over)
    u = P.rand(prob.n_dofs, 13)
    v, _ = ctx.apply_prec(u)
    Pm = prec_matrix(p2)
    assert P.relerr(Pm @ v, u) < 1e-8


def test_outer_solve_reaches_the_direct_solution(small, oracle_mod):
    """FGMRES + AL solution == sparse direct solution of the augmented system, and the
    augmented system has the same solution as the un-augmented saddle point (ii)."""
    name, prob, H = small
    ctx, p2 = ctx_for(oracle_mod, prob, H)
    rhs = P.rhs_of(ctx, p2)
    x, info = ctx.solve(rhs)
    assert info.status == 0
    AA = system_matrix(prob)
    if prob.config.kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
        # enclosed flow: pressure defined up to a constant -> compare residuals and velocity
        assert np.linalg.norm(AA @ x - rhs) <= 10 * max(prob.config.outer.tol, 1e-12 * info.initial_residual)
        xd = np.linalg.lstsq(AA, rhs, rcond=None)[0]
        n = prob.sizes[0]
        assert P.relerr(x[:n], xd[:n]) < 1e-4
    else:
        xd = np.linalg.solve(AA, rhs)
        ctl = prob.config.outer
        assert np.linalg.norm(AA @ x - rhs) <= 10 * max(ctl.tol, ctl.reduce * info.initial_residual)
        assert P.relerr(x, xd) < max(1e-4, 1e5 * ctl.reduce)  # elasticity.prm stops at a 1e-6 reduction
    if prob.config.kind == b.KIND_LAPLACE and not prob.config.aug_explicit:
        n, m = prob.Ct.shape
        K = np.block([[prob.A.toarray(), prob.Ct.toarray()], [prob.Ct.toarray().T, np.zeros((m, m))]])
        xs = np.linalg.solve(K, prob.rhs)
        assert P.relerr(x[:n], xs[:n]) < 1e-6
    assert info.outer_iterations == info.n_history - 1
    per_apply = 2 if prob.config.kind == b.KIND_ELLIPTIC_MODIFIED else 1  # A22 and A11 solves
    assert info.inner_solves == per_apply * info.outer_iterations


def test_minres_variant_converges(oracle_mod):
    prob, H = P.get("stokes2d_minres")
    ctx, p2 = ctx_for(oracle_mod, prob, H)
    rhs = P.rhs_of(ctx, p2)
    x, info = ctx.solve(rhs)
    assert info.status == 0
    true_res = np.linalg.norm(ctx.apply_system(x) - rhs)
    assert true_res < 1e-5 * info.initial_residual


def test_outer_iterations_are_mesh_independent(oracle_mod):
    """The paper's claim (invariant (v)): outer AL-FGMRES counts do not grow with refinement."""
    its = []
    for r in (5, 6, 7):
        prob = syn.immersed_laplace(r_bg=r, diagonal_inverse=True)
        H = syn.build_hierarchies(prob)
        ctx, p2 = ctx_for(oracle_mod, prob, H)
        _, info = ctx.solve(P.rhs_of(ctx, p2))
        its.append(info.outer_iterations)
    assert max(its) - min(its) <= 4 and max(its) < 30, its


def test_constraint_is_satisfied(oracle_mod):
    """|M u2 - C u|_inf ~ solver tolerance (elliptic_interface.cc:973-984, invariant (iii))."""
    prob, H = P.get("elliptic_modified")
    ctx, p2 = ctx_for(oracle_mod, prob, H)
    x, _ = ctx.solve(P.rhs_of(ctx, p2))
    n, m = prob.Ct.shape
    d = prob.M @ x[n:n + m] - prob.Ct.T @ x[:n]
    assert np.abs(d).max() < 1e-8


def test_nitsche_bcs_converges_to_the_manufactured_solution(oracle_mod):
    """nitsche_bcs through the 2x2 AL path: the FGMRES solution equals the sparse
    direct solution of the saddle-point system, approaches the manufactured solution at O(h^2), and the
    outer iteration count does not grow with refinement."""
    errs, outer = [], []
    for r in (3, 4, 5):
        prob = syn.nitsche_bcs(r)
        H = syn.build_hierarchies(prob, max_coarse=40)
        ctx = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
        x, info = ctx.solve(prob.rhs)
        n = prob.sizes[0]
        K = sp.bmat([[prob.A, prob.Ct], [prob.Ct.T, None]]).tocsc()
        xd = spla.spsolve(K, prob.rhs)
        assert info.status == 0 and P.relerr(x, xd) < 1e-4
        errs.append(np.sqrt(np.mean((x[:n] - prob.meta["u_exact"]) ** 2)))
        outer.append(info.outer_iterations)
    assert errs[1] < errs[0] / 3.5 and errs[2] < errs[1] / 3.5
    assert max(outer) <= 8


def test_nitsche_bcs_shipped_p0_multiplier_converges(oracle_mod):
    """parameters_nitsche.prm: discontinuous P0 multipliers; the alternating multiplier is in the kernel of
    the coupling matrix, FGMRES still reduces the residual of the consistent system by 1e-9."""
    prob = syn.nitsche_bcs(4, multiplier_degree=0, manufactured=False)
    alt = (-1.0) ** np.arange(prob.sizes[1])
    assert np.abs(prob.Ct @ alt).max() < 1e-14
    H = syn.build_hierarchies(prob, max_coarse=40)
    ctx = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    x, info = ctx.solve(prob.rhs)
    assert info.status == 0 and info.final_residual <= 1e-9 * info.initial_residual * 1.01
    r = prob.rhs - np.concatenate([prob.A @ x[: prob.sizes[0]] + prob.Ct @ x[prob.sizes[0]:], prob.Ct.T @ x[: prob.sizes[0]]])
    assert np.linalg.norm(r) <= 1e-8 * np.linalg.norm(prob.rhs)
