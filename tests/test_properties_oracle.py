"""The property checks of tests/properties.py on the CPU oracle (small sizes)."""
import pytest

from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from . import problems as P
from .properties import check_properties


@pytest.mark.parametrize("name", ["laplace_diag", "stokes2d_diag", "stokes2d_exact", "stokes3d_node", "elliptic_modified_diag"])
def test_properties_hold_for_the_oracle(name, oracle_mod):
    prob, H = P.get(name)
    ora = syn.setup_context(oracle_mod.OracleContext(prob.config), prob, H, oracle=True)
    info = check_properties(ora, prob)
    assert info.outer_iterations > 0
