"""Exact mass inverses in Chebyshev form (the device replacement of SparseDirectUMFPACK::vmult for mass
matrices too large for one CTA: elliptic_interface.cc:719-720, 736-737 with a co-dimension-0 multiplier space;
stokes_immersed_boundary.cc:960-962 for the pressure mass matrix).  Both forms — one fused SpMV kernel per
iteration, and the persistent kernel with the matrix staged in shared memory and one grid barrier per
iteration — against the oracle's sparse direct solves, on the small seeded cases (test knobs push them onto
this path and split them over many CTAs) and at the size of the benched elliptic problem (default selection)."""
import functools

import numpy as np
import pytest

from fictitious_domain_al_preconditioners_b200 import ALContext
from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from . import parity_log as PL
from . import problems as P
from .test_gpu_parity import TOL_SOLUTION, _self_sensitivity

pytestmark = pytest.mark.gpu

FORM = {"1": "cheb_kernels", "2": "cheb_persistent"}


@pytest.mark.parametrize("form", ["1", "2"])
@pytest.mark.parametrize("name", ["laplace_exact", "stokes2d_exact", "elliptic_modified", "elliptic_ideal", "elasticity"])
def test_chebyshev_mass_solves_match_the_direct_solves(name, form, oracle_mod, monkeypatch):
    monkeypatch.setenv("FDAL_MASS_CHEB", form)
    monkeypatch.setenv("FDAL_MASS_CHEB_MIN_ROWS", "0")  # small multiplier spaces take the Chebyshev path too
    monkeypatch.setenv("FDAL_MASS_CHEB_RPC", "16")      # 16 rows per CTA: the grid barrier is exercised
    prob, H = P.get(name)
    gpu = syn.setup_context(ALContext(prob.config), prob, H)
    ora = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    info = gpu.mass_solver_info(0)
    assert info["form"] == FORM[form], (info, gpu.last_error())
    assert 2 <= info["iterations"] <= 300 and info["verified_residual"] <= 2e-14
    PL.record(f"mass_cheb[{name},{form}]", info)
    x = P.rand(prob.sizes[-1], 4)
    PL.check("apply_winv", P.relerr(gpu.apply_winv(x), ora.apply_winv(x)), 1e-12)
    z = P.rand(prob.sizes[0], 3)
    PL.check("apply_aug", P.relerr(gpu.apply_aug(z), ora.apply_aug(z)), 1e-12)
    if prob.config.kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES) and prob.config.mp_inv_mode == b.MPINV_EXACT:
        assert gpu.mass_solver_info(1)["form"] == FORM[form]
        q = P.rand(prob.sizes[1], 5)
        PL.check("apply_mp_inv", P.relerr(gpu.apply_mp_inv(q)[0], ora.apply_mp_inv(q)[0]), 1e-12)
    X = P.rand(prob.n_dofs, 6)
    PL.check("apply_system", P.relerr(gpu.apply_system(X), ora.apply_system(X)), 1e-12)
    rhs = P.rhs_of(ora, prob)
    xg, ig = gpu.solve(rhs)
    xo, io = ora.solve(rhs)
    assert abs(ig.outer_iterations - io.outer_iterations) <= 1
    if ig.outer_iterations == io.outer_iterations:
        # the bar of test_gpu_parity.test_full_solve: 1e-10 unless the oracle's own solution moves more than that
        # under a one-ulp perturbation of the right-hand side (unstable inner CG trajectories: the elliptic cases)
        tol = max(TOL_SOLUTION, 50 * _self_sensitivity(lambda v: ora.solve(v)[0], rhs, 4))
        PL.check("solution", P.relerr(xg, xo), tol, outer=(ig.outer_iterations, io.outer_iterations))
    res = np.linalg.norm(ora.apply_system(xg) - rhs)
    assert res <= 10 * max(prob.config.outer.tol, prob.config.outer.reduce * ig.initial_residual)


def test_default_selection_and_off_switch(monkeypatch):
    """Defaults: a small multiplier space keeps the dense form, the pressure mass matrix takes the Chebyshev
    form with one kernel per iteration; FDAL_MASS_CHEB=0 keeps the Jacobi-PCG."""
    prob, H = P.get("stokes2d_exact")
    gpu = syn.setup_context(ALContext(prob.config), prob, H)
    assert gpu.mass_solver_info(0)["form"] == "dense"
    assert gpu.mass_solver_info(1)["form"] == "cheb_kernels"
    monkeypatch.setenv("FDAL_MASS_CHEB", "0")
    gpu0 = syn.setup_context(ALContext(prob.config), prob, H)
    assert gpu0.mass_solver_info(1)["form"] == "pcg_kernels"
    q = P.rand(prob.sizes[1], 5)
    PL.check("apply_mp_inv: Chebyshev vs Jacobi-PCG", P.relerr(gpu.apply_mp_inv(q)[0], gpu0.apply_mp_inv(q)[0]), 1e-13)


@functools.lru_cache(maxsize=1)
def _elliptic_cycle6():
    prob = syn.elliptic_interface(cycle=6)
    return prob, syn.build_hierarchies(prob)


@pytest.mark.parametrize("form", ["default", "2"])
def test_benched_elliptic_problem_at_full_size(form, oracle_mod, monkeypatch):
    """configs[2] as benched (cycle 6: 1 091 843 DoFs, multiplier space m = 20 609): the default selection (one
    fused kernel per Chebyshev iteration) and the persistent kernel; W^-1, the augmented operators and the block
    system against the oracle (SuperLU)."""
    if form != "default":
        monkeypatch.setenv("FDAL_MASS_CHEB", form)
    prob, H = _elliptic_cycle6()
    gpu = syn.setup_context(ALContext(prob.config), prob, H)
    ora = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    info = gpu.mass_solver_info(0)
    assert info["form"] == ("cheb_kernels" if form == "default" else FORM[form]), (info, gpu.last_error())
    PL.record(f"mass_cheb[elliptic_cycle6,{form}]", info)
    m = prob.sizes[-1]
    for seed in (1, 2):
        x = P.rand(m, seed)
        PL.check(f"apply_winv (m={m})", P.relerr(gpu.apply_winv(x), ora.apply_winv(x)), 1e-12)
    z = P.rand(prob.sizes[0], 3)
    PL.check("apply_aug A11", P.relerr(gpu.apply_aug(z), ora.apply_aug(z)), 1e-12)
    z2 = P.rand(prob.sizes[1], 4)
    PL.check("apply_aug A22", P.relerr(gpu.apply_aug(z2, which=b.AMG_A22), ora.apply_aug(z2, which=b.AMG_A22)), 1e-12)
    X = P.rand(prob.n_dofs, 6)
    PL.check("apply_system", P.relerr(gpu.apply_system(X), ora.apply_system(X)), 1e-12)
    # zero and one-hot right-hand sides: the fixed-count iteration must not produce NaNs / must stay exact
    assert np.all(gpu.apply_winv(np.zeros(m)) == 0.0)
    e = np.zeros(m)
    e[m // 2] = 1.0
    PL.check("apply_winv (one-hot)", P.relerr(gpu.apply_winv(e), ora.apply_winv(e)), 1e-12)
