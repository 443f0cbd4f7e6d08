"""Regenerates tests/golden/ref_prec_vectors.npz — the REFERENCE-DERIVED golden vectors.

Runs only in the development container (needs /root/reference).  The reference's own
augmented_lagrangian_preconditioner.h is compiled, unmodified, against the deal.II stand-in
types of oracle/ref_harness (``make -C oracle ref``) and its five ``vmult``s are executed on
seeded inputs:

  <case>/v_ref     P.vmult(u), u = problems.rand(n_dofs, 10), every LinearOperator backed by the
                   CPU oracle's operator applications (tests/problems.py CASES).  Pins the block
                   algebra of fdalo_apply_prec / fdal_apply_prec.
  <small>/v_exact  P.vmult(u), u = problems.rand(n_dofs, 13), every LinearOperator an independent
                   dense numpy operator with EXACT inverses (tests/test_oracle_known_answers.py
                   SMALL problems).  Nothing of this repository's solver code is involved, so a
                   context run with tight inner solves must reproduce it.

    python tests/golden/generate_ref_prec.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fictitious_domain_al_preconditioners_b200 import _binding as b  # noqa: E402
from fictitious_domain_al_preconditioners_b200 import synthetic as syn  # noqa: E402
from oracle import oracle, ref_prec  # noqa: E402
from tests import problems as P  # noqa: E402
from tests import test_oracle_known_answers as KA  # noqa: E402

REF_CASES = ["laplace_diag", "laplace_exact", "stokes2d_diag", "stokes2d_exact", "stokes2d_minres", "elliptic_modified",
             "elliptic_modified_diag", "elliptic_ideal", "elasticity", "stokes2d_nogd"]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_prec_vectors.npz")


def reference_with_oracle_operators(name):
    prob, H = P.get(name)
    ora = syn.setup_context(oracle.OracleContext(prob.config), prob, H, oracle=True)
    cfg = prob.config
    u = P.rand(prob.n_dofs, 10)
    return ref_prec.reference_vmult(cfg.kind, cfg.gamma, cfg.gamma_grad_div, ora.sizes, ref_prec.context_operators(ora), u)


def dense_operators(prob):
    """Exact dense operators, straight from the formulas in the reference applications."""
    cfg = prob.config
    a = KA.aug_matrices(prob)
    W, Ct = a["W"], a["Ct"]
    A11 = a["A11"]
    if cfg.kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES) and cfg.grad_div_in_operator:
        Bt = prob.Bt.toarray()
        A11 = A11 + cfg.gamma_grad_div * Bt @ np.linalg.solve(prob.Mp.toarray(), Bt.T)
    ops = {
        ref_prec.OP_AUG_INV: lambda x: np.linalg.solve(A11, x),
        ref_prec.OP_C: lambda x: Ct.T @ x,
        ref_prec.OP_CT: lambda x: Ct @ x,
        ref_prec.OP_INVW: lambda x: W @ x,
    }
    if prob.Bt is not None:
        Bt, Mp = prob.Bt.toarray(), prob.Mp.toarray()
        ops[ref_prec.OP_BT] = lambda x: Bt @ x
        ops[ref_prec.OP_MP_INV] = lambda x: np.linalg.solve(Mp, x)
    if prob.A2 is not None:
        M = a["M"]
        n0 = A11.shape[0]
        blk = np.block([[a["A11"], a["A12"]], [a["A21"], a["A22"]]])
        ops[ref_prec.OP_M] = lambda x: M @ x
        ops[ref_prec.OP_A22_INV] = lambda x: np.linalg.solve(a["A22"], x)
        ops[ref_prec.OP_AUG_INV_BLOCK] = lambda x: np.linalg.solve(blk, x)
        assert blk.shape[0] == n0 + M.shape[0]
    return ops


def reference_with_dense_operators(name):
    fac, kw = KA.SMALL[name]
    prob = fac(**kw)
    cfg = prob.config
    u = P.rand(prob.n_dofs, 13)
    return ref_prec.reference_vmult(cfg.kind, cfg.gamma, cfg.gamma_grad_div, prob.sizes, dense_operators(prob), u)


if __name__ == "__main__":
    assert ref_prec.available(), "needs /root/reference"
    out = {}
    for n in REF_CASES:
        out[f"{n}/v_ref"] = reference_with_oracle_operators(n)
    for n in KA.SMALL:
        out[f"{n}/v_exact"] = reference_with_dense_operators(n)
    np.savez_compressed(OUT, **out)
    print("wrote", len(out), "arrays,", os.path.getsize(OUT) // 1024, "KiB")
