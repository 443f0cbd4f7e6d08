"""Size-independent properties of the solve path (SURVEY 8(c) invariants) that need no second
implementation: used on the oracle at small sizes (CPU suite, which also validates this file)
and on the CUDA path at BASELINE.json's full size (GPU suite)."""
import numpy as np

from fictitious_domain_al_preconditioners_b200 import _binding as b

from . import problems as P


def check_properties(ctx, prob, scatter=lambda v: v, gather=lambda v: v, tol=1e-11, with_amg=True, op_noise=0.0):
    """ctx: a finalized ALContext / OracleContext for `prob` (vectors pass through scatter/gather
    when the context works in a renumbered space).  Returns the solve info."""
    n0 = prob.sizes[0]
    N = prob.n_dofs
    pad = lambda v: np.concatenate([v, np.zeros(N - n0)])  # noqa: E731
    x, y = P.rand(n0, 31), P.rand(n0, 32)
    xl, yl = scatter(pad(x))[:ctx.sizes[0]], scatter(pad(y))[:ctx.sizes[0]]
    # (a) the augmented operator is linear and symmetric:  y.(Ag x) == x.(Ag y)
    nrm = np.linalg.norm
    Ax, Ay = ctx.apply_aug(xl), ctx.apply_aug(yl)
    assert abs(yl @ Ax - xl @ Ay) <= tol * nrm(yl) * nrm(Ax)
    assert P.relerr(ctx.apply_aug(2.0 * xl - 3.0 * yl), 2.0 * Ax - 3.0 * Ay) < tol
    # (b) one V-cycle is a symmetric positive definite linear operator
    if with_amg:
        Bx, By = ctx.apply_amg(xl), ctx.apply_amg(yl)
        assert abs(yl @ Bx - xl @ By) <= tol * nrm(yl) * nrm(Bx)
        assert xl @ Bx > 0 and yl @ By > 0
    # (c) the block system operator is symmetric for the 2x2 / Stokes systems
    if prob.config.kind in (b.KIND_LAPLACE, b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
        X, Y = scatter(P.rand(N, 33)), scatter(P.rand(N, 34))
        AX, AY = ctx.apply_system(X), ctx.apply_system(Y)
        assert abs(Y @ AX - X @ AY) <= tol * nrm(Y) * nrm(AX)
    # (d) the inner solve meets deal.II's absolute stopping rule, measured with the operator itself
    rhs0 = xl
    sol, its = ctx.apply_aug_inv(rhs0)
    ctl = prob.config.inner
    assert 1 <= its <= ctl.max_steps
    if ctl.type == b.CONTROL_SOLVER:
        assert np.linalg.norm(rhs0 - ctx.apply_aug(sol)) <= 1.05 * ctl.tol
    # (e) the outer solve: the TRUE residual of the returned solution meets the outer control,
    # the residual history decreases to it, and the reported counts are consistent
    rhs = scatter(prob.rhs)
    if prob.augment_rhs:
        rhs = ctx.augment_rhs(rhs)
    sol, info = ctx.solve(rhs)
    assert info.status == 0
    true_res = np.linalg.norm(ctx.apply_system(sol) - rhs)
    oc = prob.config.outer
    # op_noise: level at which the operator itself is reproducible (nested inexact solves)
    assert true_res <= 10.0 * max(oc.tol, oc.reduce * info.initial_residual, op_noise)
    h = info.history()
    assert h[-1] <= h[0] and info.outer_iterations == len(h) - 1
    assert info.inner_solves >= info.outer_iterations
    # (f) the multiplier enforces the coupling constraint: C u = g (immersed_laplace / Stokes)
    if prob.augment_rhs:
        full = gather(sol)
        n, m = prob.Ct.shape
        g = prob.rhs[-m:]
        assert np.linalg.norm(prob.Ct.T @ full[:n] - g) <= 10.0 * max(oc.tol, oc.reduce * info.initial_residual, op_noise)
    return info
