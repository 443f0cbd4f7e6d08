"""The gamma parameter study of elliptic_interface (elliptic_interface.cc:1086-1128; utilities.h:333-346) on the
CPU oracle: linspace, first-minimum rule, the dependence of the outer count on gamma that makes the study worth
running.  The same study through the CUDA library: tests/test_gpu_variants.py."""
import numpy as np
import pytest

from fictitious_domain_al_preconditioners_b200 import parameter_study as ps
from fictitious_domain_al_preconditioners_b200 import synthetic as syn


def test_linspace_is_the_reference_helper():
    assert ps.linspace(1e-3, 1.0, 100)[0] == 1e-3
    assert ps.linspace(1e-3, 1.0, 100)[-1] == pytest.approx(1.0, rel=1e-15)
    assert np.allclose(ps.linspace(0.0, 2.0, 5), [0.0, 0.5, 1.0, 1.5, 2.0])
    with pytest.raises(ValueError):
        ps.linspace(1.0, 1.0, 4)  # AssertThrow(start < end)


def test_first_minimum_wins():
    s = ps.ParameterStudy(gammas=[0.1, 0.2, 0.3, 0.4], outer_iterations=[9, 7, 7, 8])
    assert s.min_index == 1 and s.best_gamma == 0.2


def test_gamma_study_on_the_oracle(oracle_mod):
    gammas = ps.linspace(1e-3, 10.0, 4)
    study = ps.gamma_parameter_study(
        lambda g: syn.elliptic_interface(cycle=1, gamma_fluid=g, gamma_solid=g),
        gammas, make_context=lambda cfg: oracle_mod.OracleContext(cfg),
        build_hierarchies=lambda p: syn.build_hierarchies(p, max_coarse=300))
    assert study.gammas == gammas and len(study.outer_iterations) == 4
    assert all(1 <= k <= 1000 for k in study.outer_iterations)
    # gamma matters: the smallest sampled value needs more outer iterations than the best one
    assert study.outer_iterations[0] > min(study.outer_iterations)
    assert study.best_gamma == gammas[int(np.argmin(study.outer_iterations))]
