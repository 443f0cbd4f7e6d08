"""Host-side mirror of the reference operator API for the AL path.

Same names, argument meaning and error behaviour as
``augmented_lagrangian_preconditioner.h`` and the ``LinearOperator`` expressions in the
three ``solve()`` functions, so that code (and tests) written against the reference
read the same here:

    ops   = Operators(ctx)                       # handles on one fdal_ctx
    Aug   = ops.Aug ; Aug_inv = ops.Aug_inv      # K + gamma*Ct*invW*C , inverse_operator(Aug, CG, AMG)
    P     = BlockPreconditionerAugmentedLagrangian(Aug_inv, ops.C, ops.Ct, ops.invW, gamma)
    SolverFGMRES(control).solve(ops.AA, x, b, P) # -> fdal_solve, whole solve on the device

Every ``LinearOperator.vmult`` is ONE call through the C ABI into the CUDA library
(per-vmult "parity mode": host<->device copy per call).  The preconditioner classes
evaluate the reference's literal block algebra with those operators — the host only
adds/scales the small glue vectors exactly where deal.II's ``BlockVector`` expressions
would — while ``vmult_fused`` / ``SolverFGMRES.solve`` use the device-resident fused
path (``fdal_apply_prec`` / ``fdal_solve``), which is the product.
"""
from __future__ import annotations

import numpy as np

from . import _binding as b
from .context import ALContext, NoConvergence, SolverControl  # noqa: F401


class BlockVector:
    """dealii::BlockVector<double>: blocks are views of one contiguous array."""

    def __init__(self, sizes, data=None):
        self.sizes = tuple(int(s) for s in sizes)
        self.data = np.zeros(sum(self.sizes)) if data is None else np.ascontiguousarray(data, dtype=np.float64)
        assert self.data.size == sum(self.sizes)
        self._off = np.concatenate([[0], np.cumsum(self.sizes)])

    def block(self, i):
        return self.data[self._off[i]: self._off[i + 1]]

    def n_blocks(self):
        return len(self.sizes)

    def copy(self):
        return BlockVector(self.sizes, self.data.copy())


class LinearOperator:
    """dealii::LinearOperator<Vector<double>>: a vmult (and optional Tvmult) closure."""

    def __init__(self, vmult, n_rows, n_cols, tvmult=None, ctx=None, tag=None):
        self._vmult, self._tvmult = vmult, tvmult
        self.n_rows, self.n_cols = n_rows, n_cols
        self.ctx, self.tag = ctx, tag
        self.is_null_operator = vmult is None

    def vmult(self, dst, src):
        dst[...] = self._vmult(np.asarray(src))

    def Tvmult(self, dst, src):
        dst[...] = self._tvmult(np.asarray(src))

    def __mul__(self, src):  # op * vector
        return self._vmult(np.asarray(src))

    __matmul__ = __mul__


def transpose_operator(op: LinearOperator) -> LinearOperator:
    """transpose_operator(linear_operator(Ct)) — immersed_laplace.cc:641."""
    return LinearOperator(op._tvmult, op.n_cols, op.n_rows, tvmult=op._vmult, ctx=op.ctx,
                          tag=("T", op.tag))


def linear_operator(ctx: ALContext, matrix_id: int, shape) -> LinearOperator:
    """linear_operator(SparseMatrix) — vmult / Tvmult of an exported CSR block (fdal_spmv)."""
    nr, nc = shape
    return LinearOperator(
        lambda x: ctx.spmv(matrix_id, x, transpose=False, n_out=nr),
        nr, nc,
        tvmult=lambda x: ctx.spmv(matrix_id, x, transpose=True, n_out=nc),
        ctx=ctx, tag=("mat", matrix_id),
    )


class Operators:
    """The operator objects the three reference ``solve()`` functions build, bound to one context."""

    def __init__(self, ctx: ALContext):
        self.ctx = ctx
        cfg = ctx.config
        self.sizes = ctx.sizes
        n0 = self.sizes[0]
        m = self.sizes[-1]
        self.K = self.A = linear_operator(ctx, b.MAT_A, (n0, n0))
        self.Ct = linear_operator(ctx, b.MAT_CT, (n0, m))
        self.C = transpose_operator(self.Ct)
        self.invW = LinearOperator(ctx.apply_winv, m, m, tvmult=ctx.apply_winv, ctx=ctx, tag="invW")
        if b.MAT_M in ctx.matrices:
            self.M = linear_operator(ctx, b.MAT_M, (m, m))
        # Aug = K + gamma * Ct * invW * C  (one fused call) and its inverse
        self.Aug = LinearOperator(lambda x: ctx.apply_aug(x, b.AMG_A11), n0, n0, ctx=ctx, tag="Aug")
        self.Aug_inv = LinearOperator(lambda x: ctx.apply_aug_inv(x, b.AMG_A11)[0], n0, n0, ctx=ctx, tag="Aug_inv")
        self.A11_aug, self.A11_aug_inv = self.Aug, self.Aug_inv
        if cfg.kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
            n_p = self.sizes[1]
            self.Bt = linear_operator(ctx, b.MAT_BT, (n0, n_p))
            self.B = transpose_operator(self.Bt)
            self.Mp = linear_operator(ctx, b.MAT_MP, (n_p, n_p))
            self.Mp_inv = LinearOperator(lambda x: ctx.apply_mp_inv(x)[0], n_p, n_p, ctx=ctx, tag="Mp_inv")
        if cfg.kind in (b.KIND_ELLIPTIC_IDEAL, b.KIND_ELLIPTIC_MODIFIED):
            self.A_omega2 = linear_operator(ctx, b.MAT_A2, (m, m))
            self.A22_aug = LinearOperator(lambda x: ctx.apply_aug(x, b.AMG_A22), m, m, ctx=ctx, tag="A22_aug")
            self.A22_aug_inv = LinearOperator(lambda x: ctx.apply_aug_inv(x, b.AMG_A22)[0], m, m, ctx=ctx,
                                              tag="A22_aug_inv")
        # AA / system_operator: block_operator<...>
        self.AA = self.system_operator = BlockOperator(ctx)


class BlockOperator:
    """block_operator<2,2 / 3,3, BlockVector<double>> of the augmented system."""

    def __init__(self, ctx):
        self.ctx = ctx

    def vmult(self, dst: BlockVector, src: BlockVector):
        dst.data[...] = self.ctx.apply_system(src.data)


class _PrecBase:
    def __init__(self, ctx):
        self.ctx = ctx

    def vmult_fused(self, v: BlockVector, u: BlockVector):
        """Device-resident evaluation (fdal_apply_prec): what fdal_solve uses."""
        v.data[...], self.last_inner_iterations = self.ctx.apply_prec(u.data)


class BlockPreconditionerAugmentedLagrangian(_PrecBase):
    """augmented_lagrangian_preconditioner.h:14-42."""

    def __init__(self, Aug_inv_, C_, Ct_, invW_, gamma_=1e2):
        super().__init__(Aug_inv_.ctx)
        self.Aug_inv, self.C, self.Ct, self.invW, self.gamma = Aug_inv_, C_, Ct_, invW_, gamma_

    def vmult(self, v: BlockVector, u: BlockVector):
        v.block(1)[...] = -self.gamma * (self.invW * u.block(1))
        v.block(0)[...] = self.Aug_inv * (u.block(0) - self.Ct * v.block(1))


class BlockPreconditionerAugmentedLagrangianStokes(_PrecBase):
    """augmented_lagrangian_preconditioner.h:44-79."""

    def __init__(self, Aug_inv_, Bt_, Ct_, invW_, Mp_inv_, gamma_, gamma_grad_div_):
        super().__init__(Aug_inv_.ctx)
        self.Aug_inv, self.Bt, self.Ct, self.invW, self.Mp_inv = Aug_inv_, Bt_, Ct_, invW_, Mp_inv_
        self.gamma, self.gamma_grad_div = gamma_, gamma_grad_div_

    def vmult(self, v: BlockVector, u: BlockVector):
        v.block(2)[...] = -self.gamma * (self.invW * u.block(2))
        v.block(1)[...] = -self.gamma_grad_div * (self.Mp_inv * u.block(1))
        v.block(0)[...] = self.Aug_inv * (u.block(0) - self.Bt * v.block(1) - self.Ct * v.block(2))


class BlockPreconditionerAugmentedLagrangianDiagonal(_PrecBase):
    """augmented_lagrangian_preconditioner.h:81-110 (SPD variant for MinRes)."""

    def __init__(self, Aug_inv_, invW_, Mp_inv_, gamma_, gamma_grad_div_):
        super().__init__(Aug_inv_.ctx)
        self.Aug_inv, self.invW, self.Mp_inv = Aug_inv_, invW_, Mp_inv_
        self.gamma, self.gamma_grad_div = gamma_, gamma_grad_div_

    def vmult(self, v: BlockVector, u: BlockVector):
        v.block(2)[...] = self.gamma * (self.invW * u.block(2))
        v.block(1)[...] = self.gamma_grad_div * (self.Mp_inv * u.block(1))
        v.block(0)[...] = self.Aug_inv * u.block(0)


class BlockTriangularALPreconditionerModified(_PrecBase):
    """EllipticInterfacePreconditioners::BlockTriangularALPreconditionerModified
    (augmented_lagrangian_preconditioner.h:168-238)."""

    def __init__(self, C_, M_, invW_, gamma_, A11_inv_, A22_inv_):
        super().__init__(A11_inv_.ctx)
        self.C, self.Ct, self.M, self.invW = C_, transpose_operator(C_), M_, invW_
        self.gamma, self.A11_inv, self.A22_inv = gamma_, A11_inv_, A22_inv_

    def vmult(self, dst: BlockVector, src: BlockVector):
        assert src.n_blocks() == 3 and dst.n_blocks() == 3
        u, u2, lam = src.block(0), src.block(1), src.block(2)
        dst.block(2)[...] = -self.gamma * (self.invW * lam)
        dst.block(1)[...] = self.A22_inv * (u2 + self.M * dst.block(2))
        dst.block(0)[...] = self.A11_inv * (
            u + self.gamma * (self.Ct * (self.invW * (self.M * dst.block(1)))) - self.Ct * dst.block(2))


class BlockTriangularALPreconditioner(_PrecBase):
    """EllipticInterfacePreconditioners::BlockTriangularALPreconditioner ("ideal",
    augmented_lagrangian_preconditioner.h:115-164): Aug_inv acts on the first two
    blocks together, so only the fused evaluation exists here."""

    def __init__(self, ctx, C_, M_, invW_, gamma_):
        super().__init__(ctx)
        self.C, self.M, self.invW, self.gamma = C_, M_, invW_, gamma_

    def vmult(self, v: BlockVector, u: BlockVector):
        self.vmult_fused(v, u)


class SolverFGMRES:
    """SolverFGMRES<BlockVector<double>>: solve(A, x, b, P) runs entirely on the device
    (fdal_solve).  Raises NoConvergence like dealii::SolverControl::NoConvergence."""

    def __init__(self, control: SolverControl | None = None, max_basis_size: int | None = None):
        self.control, self.max_basis_size = control, max_basis_size
        self.last_info = None

    def solve(self, A, x: BlockVector, rhs: BlockVector, P):
        ctx = A.ctx
        assert P.ctx is ctx, "operator and preconditioner must live on the same context"
        sol, info = ctx.solve(rhs.data, x0=x.data)
        x.data[...] = sol
        self.last_info = info
        return info


SolverMinRes = SolverFGMRES  # the context's kind selects MinRes (FDAL_KIND_STOKES_DIAG_MINRES)
