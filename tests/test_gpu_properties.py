"""Size-independent properties on the CUDA path at BASELINE.json's full size (configs[1]: the
2-D Stokes immersed-boundary problem at ~1 M DoFs, BSR path) and on a 3-D Stokes problem —
sizes where running the oracle would take minutes."""
import copy

import pytest

from fictitious_domain_al_preconditioners_b200 import ALContext
from fictitious_domain_al_preconditioners_b200 import partition as part
from fictitious_domain_al_preconditioners_b200 import synthetic as syn

from .properties import check_properties

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kw", [
    pytest.param(dict(dim=2, nel=320, diagonal_mass=False), id="stokes2d_1M_as_shipped"),
    pytest.param(dict(dim=3, nel=16), id="stokes3d_nel16")])
def test_properties_at_full_size(kw):
    prob = syn.stokes_immersed_boundary(numbering="node", **kw)
    H = syn.build_hierarchies(prob)
    lp = part.distribute_problem(prob, H, 0, 1)
    cfg = copy.deepcopy(prob.config)
    cfg.block_size = lp.block_size
    gpu = part.setup_local_context(ALContext(cfg), lp)
    info = check_properties(gpu, prob, scatter=lp.scatter, gather=lambda v: lp.gather([v]))
    assert 5 <= info.outer_iterations <= 40 and info.kernel_launches > 0
