"""Dim-blocked (BSR) storage of the velocity block and of the finest AMG operator: the
node-interleaved renumbering + BSR kernels must reproduce the scalar-CSR oracle."""
import copy

import numpy as np
import pytest

from fictitious_domain_al_preconditioners_b200 import ALContext
from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import partition as part
from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from . import parity_log as PL
from . import problems as P
from .test_gpu_parity import _self_sensitivity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["stokes2d_diag", "stokes2d_exact", "stokes3d_diag", "stokes3d_node", "stokes2d_node", "elasticity", "elasticity_diag"])
def test_bsr_path_matches_oracle(name, oracle_mod):
    prob, H = P.get(name)
    lp = part.distribute_problem(prob, H, 0, 1)
    assert lp.block_size in (2, 3)
    cfg = copy.deepcopy(prob.config)
    cfg.block_size = lp.block_size
    gpu = part.setup_local_context(ALContext(cfg), lp)
    ora = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    X = P.rand(prob.n_dofs, 5)
    y = lp.gather([gpu.apply_system(lp.scatter(X))])
    PL.check("BSR apply_system", P.relerr(y, ora.apply_system(X)), 1e-12)
    n0 = prob.sizes[0]
    r = P.rand(n0, 7)
    rpad = np.concatenate([r, np.zeros(prob.n_dofs - n0)])
    z = gpu.apply_amg(lp.scatter(rpad)[:n0])
    zfull = lp.gather([np.concatenate([z, np.zeros(prob.n_dofs - n0)])])[:n0]
    PL.check("BSR apply_amg", P.relerr(zfull, ora.apply_amg(r)), 1e-12)
    xa = gpu.apply_aug(lp.scatter(rpad)[:n0])
    xafull = lp.gather([np.concatenate([xa, np.zeros(prob.n_dofs - n0)])])[:n0]
    PL.check("BSR apply_aug", P.relerr(xafull, ora.apply_aug(r)), 1e-12)
    u = P.rand(prob.n_dofs, 10)
    v, its = gpu.apply_prec(lp.scatter(u))
    vo, ito = ora.apply_prec(u)
    assert tuple(its) == tuple(ito)
    tol = max(1e-10, 50 * _self_sensitivity(lambda w: ora.apply_prec(w)[0], u, 2))
    PL.check("BSR apply_prec", P.relerr(lp.gather([v]), vo), tol, its_gpu=list(its), its_oracle=list(ito))
    rhs = P.rhs_of(ora, prob)
    xg, ig = gpu.solve(lp.scatter(rhs))
    xo, io = ora.solve(rhs)
    assert abs(ig.outer_iterations - io.outer_iterations) <= 1
    if ig.outer_iterations == io.outer_iterations:
        tol = max(1e-10, 50 * _self_sensitivity(lambda w: ora.solve(w)[0], rhs, 4))
        PL.check("BSR solve: solution", P.relerr(lp.gather([xg]), xo), tol, outer_gpu=int(ig.outer_iterations),
                 outer_oracle=int(io.outer_iterations))
    # the BSR kernels were the ones that ran
    ms, by, nl = gpu.time_kernel(b.TIME_SPMV_A, 0, warmup=1, reps=2, flush_l2=False)
    bs = lp.block_size
    assert by < 12.0 * prob.A.nnz + 24.0 * n0  # fewer bytes than scalar CSR => BSR storage active
