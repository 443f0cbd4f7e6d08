"""The reference-side binding ``include/fdal_dealii.h``, compiled and run.

deal.II / Trilinos are absent, so the adapter is compiled against the stand-in deal.II types of
oracle/ref_harness/dealii_stub — TOGETHER WITH the reference's own preconditioner classes
(augmented_lagrangian_preconditioner.h, unmodified) — and bound to the CPU oracle
(oracle/ref_harness/adapter_check.cc).  What a patched reference application would do is done here:
the reference class is constructed from the adapter's LinearOperators and applied; the fused
``ALPreconditioner`` and ``solve`` wrappers are called; failures must surface as
``SolverControl::NoConvergence``.  The CUDA flavour of the same library is exercised on the GPU in
tests/test_zz_gpu_unverified.py.
"""
import copy

import numpy as np
import pytest

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn
from fictitious_domain_al_preconditioners_b200.context import SolverControl
from oracle import adapter_check as ac

from . import problems as P

pytestmark = pytest.mark.skipif(not ac.available("oracle"), reason="/root/reference is not mounted here")

NAMES = ["laplace_diag", "laplace_exact", "laplace_opform", "stokes2d_diag", "stokes2d_exact", "stokes2d_minres",
         "stokes3d_diag", "elliptic_modified", "elliptic_modified_diag", "elliptic_ideal", "elasticity", "nitsche_p1"]


def _oracle(oracle_mod, name, **over):
    prob, H = P.get(name)
    cfg = copy.deepcopy(prob.config)
    for k, v in over.items():
        setattr(cfg, k, v)
    p2 = copy.copy(prob)
    p2.config = cfg
    return p2, syn.setup_context(oracle_mod.OracleContext(cfg), p2, H, oracle=True)


@pytest.mark.parametrize("name", NAMES)
def test_reference_class_on_adapter_operators_reproduces_apply_prec(name, oracle_mod):
    lib = ac.load("oracle")
    prob, ctx = _oracle(oracle_mod, name)
    u = P.rand(prob.n_dofs, 10)
    v_ref, st = ac.reference_vmult(lib, ctx, u)
    assert st == 0
    v, its = ctx.apply_prec(u)
    assert P.relerr(v_ref, v) < 1e-13
    v_al, its_al, st = ac.al_vmult(lib, ctx, u)
    assert st == 0 and tuple(its_al) == tuple(its) and np.array_equal(v_al, v)


@pytest.mark.parametrize("name", ["laplace_diag", "stokes2d_exact", "elliptic_modified", "elliptic_ideal"])
def test_adapter_solve_is_the_c_abi_solve(name, oracle_mod):
    lib = ac.load("oracle")
    prob, ctx = _oracle(oracle_mod, name)
    rhs = P.rhs_of(ctx, prob)
    x, info = ctx.solve(rhs)
    xa, infa, st = ac.solve(lib, ctx, rhs)
    assert st == 0 and np.array_equal(xa, x)
    assert infa.outer_iterations == info.outer_iterations and infa.inner_iterations == info.inner_iterations


def test_failures_surface_as_no_convergence(oracle_mod):
    lib = ac.load("oracle")
    # one inner CG step cannot reach 1e-10: SolverCG would throw NoConvergence inside Aug_inv.vmult
    prob, ctx = _oracle(oracle_mod, "laplace_diag", inner=SolverControl(1, 1e-10))
    u = P.rand(prob.n_dofs, 10)
    assert ac.reference_vmult(lib, ctx, u)[1] == 1
    assert ac.al_vmult(lib, ctx, u)[2] == 1
    assert ac.solve(lib, ctx, P.rhs_of(ctx, prob))[2] == 1
    # outer failure
    prob, ctx = _oracle(oracle_mod, "laplace_diag", outer=SolverControl(2, 1e-14))
    x, info, st = ac.solve(lib, ctx, P.rhs_of(ctx, prob))
    assert st == 1 and info.status == b.ERR_OUTER_NO_CONVERGENCE


def test_export_csr_from_a_dealii_sparse_matrix(oracle_mod):
    """Context populated through fdal_dealii::export_csr (row iterators of dealii::SparseMatrix)
    behaves like one populated through fdal_set_csr directly."""
    lib = ac.load("oracle")
    prob, H = P.get("stokes2d_diag")
    ref = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    ctx = oracle_mod.OracleContext(prob.config)
    for mid, A in ((b.MAT_A, prob.A), (b.MAT_CT, prob.Ct), (b.MAT_BT, prob.Bt), (b.MAT_MP, prob.Mp), (b.MAT_M, prob.M)):
        assert ac.export_csr(lib, ctx, mid, A) == 0
    ctx.set_diag(b.DIAG_W_INV, prob.winv_diag)
    for which, Hh in H.items():
        ctx.set_amg(which, Hh)
    ctx.finalize()
    X = P.rand(prob.n_dofs, 5)
    assert np.array_equal(ctx.apply_system(X), ref.apply_system(X))
    # a malformed matrix is refused with the library's message, as an exception on the C++ side
    ctx2 = oracle_mod.OracleContext(prob.config)
    assert ac.export_csr(lib, ctx2, 999, prob.M) == 2


def test_to_control_maps_the_solver_control_family():
    lib = ac.load("oracle")
    c = ac.to_control(lib, b.CONTROL_SOLVER, 100, 1e-2)
    assert (c.type, c.max_steps, c.tol) == (b.CONTROL_SOLVER, 100, 1e-2)
    c = ac.to_control(lib, b.CONTROL_REDUCTION, 1000, 1e-12, 1e-9)
    assert (c.type, c.max_steps, c.tol, c.reduce) == (b.CONTROL_REDUCTION, 1000, 1e-12, 1e-9)
    c = ac.to_control(lib, b.CONTROL_ITERATION_NUMBER, 7, 1e-20)
    assert (c.type, c.max_steps) == (b.CONTROL_ITERATION_NUMBER, 7)


@pytest.mark.parametrize("name", ["laplace_diag", "stokes2d_diag", "elliptic_modified"])
def test_export_amg_through_the_ml_standins(name, oracle_mod):
    """fdal_dealii::export_amg (ML_Epetra::MultiLevelPreconditioner::GetML -> Amat/Pmat/Rmat ->
    ML_Operator2EpetraCrsMatrix -> ExtractMyRowView -> fdal_amg_set_level) compiled against the ML / Epetra
    stand-ins of oracle/ref_harness/trilinos_stub and run on a hierarchy laid out the way ML lays it out:
    the context it fills applies the same V-cycle, bit for bit, as one filled through fdal_amg_set_level."""
    lib = ac.load("oracle")
    prob, H = P.get(name)
    ref = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    ctx = oracle_mod.OracleContext(prob.config)
    for mid, A in ((b.MAT_A, prob.A), (b.MAT_A2, prob.A2), (b.MAT_CT, prob.Ct), (b.MAT_BT, prob.Bt), (b.MAT_MP, prob.Mp),
                   (b.MAT_M, prob.M)):
        if A is not None:
            ctx.set_csr(mid, A)
    if prob.winv_diag is not None:
        ctx.set_diag(b.DIAG_W_INV, prob.winv_diag)
    if prob.config.winv_mode != b.WINV_DIAG:
        ctx.set_lu(0, prob.M)
    for which, Hh in H.items():
        assert ac.export_amg(lib, ctx, which, Hh) == 0
    ctx.finalize()
    r = P.rand(prob.sizes[0], 7)
    assert np.array_equal(ctx.apply_amg(r), ref.apply_amg(r))
    u = P.rand(prob.n_dofs, 10)
    assert np.array_equal(ctx.apply_prec(u)[0], ref.apply_prec(u)[0])
