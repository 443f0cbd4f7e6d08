"""Pins the oracle (and the Python mirror) against the REFERENCE CODE ITSELF.

The reference applications need deal.II/Trilinos and cannot be built here, but the header that
holds the five AL preconditioner classes is self-contained apart from deal.II's vector /
LinearOperator types.  ``oracle/ref_prec.py`` compiles that header — unmodified, from
/root/reference — against stand-in types and runs its ``vmult``s:

* live tests (development container only, skipped where /root/reference is absent) execute the
  reference vmult with the oracle's operators and demand agreement with ``fdalo_apply_prec``;
* the committed vectors of tests/golden/ref_prec_vectors.npz (made by
  tests/golden/generate_ref_prec.py from the same reference code) are checked everywhere.

Tolerances: the block algebra is a handful of AXPYs around the same inner solves, so the oracle
is required to match to 1e-13 (it is bit-identical today); `v_exact` was produced with exact
dense inverses, a context with 1e-13 inner solves must match it to 1e-8.
"""
import os

import numpy as np
import pytest

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import operators as op
from fictitious_domain_al_preconditioners_b200 import synthetic as syn
from fictitious_domain_al_preconditioners_b200.context import SolverControl
from oracle import ref_prec

from . import problems as P
from . import test_oracle_known_answers as KA
from .golden import generate_ref_prec as G

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_prec_vectors.npz"))
needs_reference = pytest.mark.skipif(not ref_prec.available(), reason="/root/reference is not mounted here")

EXPECTED_CALLS = {
    # the operator applications each reference vmult performs, in order (header lines 32-33, 67-69,
    # 100-102, 135-152, 225-228)
    b.KIND_LAPLACE: ["invW", "Ct", "Aug_inv"],
    b.KIND_STOKES: ["invW", "Mp_inv", "Bt", "Ct", "Aug_inv"],
    b.KIND_STOKES_DIAG_MINRES: ["invW", "Mp_inv", "Aug_inv"],
    b.KIND_ELLIPTIC_IDEAL: ["invW", "Ct", "M", "Aug_inv(block)"],
    b.KIND_ELLIPTIC_MODIFIED: ["invW", "M", "A22_inv", "M", "invW", "Ct", "Ct", "Aug_inv"],
}


def _oracle(oracle_mod, name):
    prob, H = P.get(name)
    return prob, syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)


@needs_reference
def test_harness_is_built_from_the_unmodified_reference_header():
    lib = ref_prec._load()
    assert lib.ref_prec_source().decode() == ref_prec.REFERENCE_HEADER
    # nothing of the reference is vendored: the harness only #includes it
    src = open(os.path.join(os.path.dirname(ref_prec.__file__), "ref_harness", "ref_prec_harness.cc")).read()
    assert "#include <augmented_lagrangian_preconditioner.h>" in src and "class BlockPreconditioner" not in src


@needs_reference
@pytest.mark.parametrize("name", list(P.CASES) + list(P.EXTRA_CASES))
def test_oracle_preconditioner_equals_the_reference_vmult(name, oracle_mod):
    prob, ora = _oracle(oracle_mod, name)
    cfg = prob.config
    u = P.rand(prob.n_dofs, 10)
    trace = []
    v_ref = ref_prec.reference_vmult(cfg.kind, cfg.gamma, cfg.gamma_grad_div, ora.sizes, ref_prec.context_operators(ora),
                                     u, trace)
    assert [ref_prec.OP_NAMES[i] for i in trace] == EXPECTED_CALLS[cfg.kind]
    v, _ = ora.apply_prec(u)
    assert P.relerr(v, v_ref) < 1e-13


@needs_reference
@pytest.mark.parametrize("name", ["laplace_diag", "stokes2d_diag", "stokes2d_minres", "elliptic_ideal", "elliptic_modified"])
def test_pinning_has_teeth(name, oracle_mod):
    """A sign or scaling slip in the oracle would be seen: the reference vmult with a perturbed
    gamma, or with one operator negated, is far from the oracle's answer."""
    prob, ora = _oracle(oracle_mod, name)
    cfg = prob.config
    u = P.rand(prob.n_dofs, 10)
    v, _ = ora.apply_prec(u)
    ops = ref_prec.context_operators(ora)
    wrong_gamma = ref_prec.reference_vmult(cfg.kind, -cfg.gamma, cfg.gamma_grad_div, ora.sizes, ops, u)
    assert P.relerr(v, wrong_gamma) > 1e-3
    flipped = dict(ops)
    flipped[ref_prec.OP_INVW] = lambda x: -ops[ref_prec.OP_INVW](x)
    assert P.relerr(v, ref_prec.reference_vmult(cfg.kind, cfg.gamma, cfg.gamma_grad_div, ora.sizes, flipped, u)) > 1e-3


def _mirror(ctx, cfg):
    o = op.Operators(ctx)
    k = cfg.kind
    if k == b.KIND_LAPLACE:
        return op.BlockPreconditionerAugmentedLagrangian(o.Aug_inv, o.C, o.Ct, o.invW, cfg.gamma)
    if k == b.KIND_STOKES:
        return op.BlockPreconditionerAugmentedLagrangianStokes(o.Aug_inv, o.Bt, o.Ct, o.invW, o.Mp_inv, cfg.gamma,
                                                               cfg.gamma_grad_div)
    if k == b.KIND_STOKES_DIAG_MINRES:
        return op.BlockPreconditionerAugmentedLagrangianDiagonal(o.Aug_inv, o.invW, o.Mp_inv, cfg.gamma,
                                                                 cfg.gamma_grad_div)
    if k == b.KIND_ELLIPTIC_MODIFIED:
        return op.BlockTriangularALPreconditionerModified(o.C, o.M, o.invW, cfg.gamma, o.A11_aug_inv, o.A22_aug_inv)
    return op.BlockTriangularALPreconditioner(ctx, o.C, o.M, o.invW, cfg.gamma)


@needs_reference
@pytest.mark.parametrize("name", ["laplace_exact", "stokes2d_exact", "stokes2d_minres", "elliptic_modified", "elliptic_ideal",
                                  "elasticity"])
def test_python_mirror_classes_equal_the_reference_classes(name, oracle_mod):
    """operators.py mirrors the header class by class (same constructor arguments, same vmult):
    composed operator by operator it gives what the reference classes give."""
    prob, ora = _oracle(oracle_mod, name)
    cfg = prob.config
    u = op.BlockVector(ora.sizes, P.rand(prob.n_dofs, 10))
    v = op.BlockVector(ora.sizes)
    _mirror(ora, cfg).vmult(v, u)
    v_ref = ref_prec.reference_vmult(cfg.kind, cfg.gamma, cfg.gamma_grad_div, ora.sizes, ref_prec.context_operators(ora),
                                     u.data)
    assert P.relerr(v.data, v_ref) < 1e-13


# ---- committed reference-derived vectors: run everywhere -----------------------------------------
@pytest.mark.parametrize("name", G.REF_CASES)
def test_oracle_reproduces_the_committed_reference_vectors(name, oracle_mod):
    prob, ora = _oracle(oracle_mod, name)
    v, _ = ora.apply_prec(P.rand(prob.n_dofs, 10))
    assert P.relerr(v, GOLD[f"{name}/v_ref"]) < 1e-12


def tight_context(make_ctx, name):
    fac, kw = KA.SMALL[name]
    prob = fac(**kw)
    H = syn.build_hierarchies(prob, max_coarse=40) if prob.amg_matrix else {}
    over = dict(inner=SolverControl(2000, 1e-13))
    if prob.config.kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
        over["mass"] = SolverControl(500, 1e-14)
    return prob, make_ctx(prob, H, over)


@pytest.mark.parametrize("name", list(KA.SMALL))
def test_oracle_with_tight_inner_solves_matches_exact_reference_vectors(name, oracle_mod):
    """v_exact involves none of this repository's solver code (dense numpy operators inside the
    reference's vmult)."""
    prob, ctx = tight_context(lambda p, H, over: KA.ctx_for(oracle_mod, p, H, **over)[0], name)
    v, _ = ctx.apply_prec(P.rand(prob.n_dofs, 13))
    assert P.relerr(v, GOLD[f"{name}/v_exact"]) < 1e-8


@needs_reference
def test_committed_vectors_are_what_the_reference_produces_today():
    for name in ("laplace_diag", "stokes2d_minres", "elliptic_ideal"):
        assert np.array_equal(G.reference_with_oracle_operators(name), GOLD[f"{name}/v_ref"])
    for name in ("stokes2d_exact", "elliptic_modified", "elasticity"):
        assert P.relerr(G.reference_with_dense_operators(name), GOLD[f"{name}/v_exact"]) < 1e-12
