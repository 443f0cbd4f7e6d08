"""Regenerates tests/golden/oracle_regression.npz.

The reference-derived vectors live in ref_prec_vectors.npz (generate_ref_prec.py); the
reference ships no tests and its applications cannot be built here, so
there is no reference data for whole solves.  The fixtures of THIS file are
produced by the CPU ORACLE itself on the seeded synthetic problems of tests/problems.py
and only pin the oracle against accidental change (iteration counts, residual histories,
solution checksums); they are not evidence of parity with deal.II / Trilinos.

    python tests/golden/generate.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fictitious_domain_al_preconditioners_b200 import synthetic as syn  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import problems as P  # noqa: E402

CASES = ["laplace_diag", "laplace_exact", "stokes2d_diag", "stokes2d_exact", "stokes3d_diag", "elliptic_modified_diag",
         "elliptic_ideal", "elasticity"]


def record(name):
    prob, H = P.get(name)
    ora = syn.setup_context(oracle.OracleContext(prob.config), prob, H, oracle=True)
    rhs = P.rhs_of(ora, prob)
    x, info = ora.solve(rhs)
    u = P.rand(prob.n_dofs, 10)
    v, its = ora.apply_prec(u)
    X = P.rand(prob.n_dofs, 5)
    y = ora.apply_system(X)
    probe = np.linspace(0, prob.n_dofs - 1, 16).astype(int)
    return {
        f"{name}/outer": info.outer_iterations,
        f"{name}/inner": info.inner_iterations,
        f"{name}/inner_a22": info.inner_iterations_a22,
        f"{name}/history": info.history()[:8],
        f"{name}/x_norm": np.linalg.norm(x),
        f"{name}/x_probe": x[probe],
        f"{name}/prec_its": np.array(its),
        f"{name}/prec_probe": v[probe],
        f"{name}/system_probe": y[probe],
    }


if __name__ == "__main__":
    out = {}
    for n in CASES:
        out.update(record(n))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression.npz"), **out)
    print("wrote", len(out), "arrays")
