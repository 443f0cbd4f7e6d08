"""Setup-phase work on the device: the operator-form AL term
gamma * sum_q phi_i(x_q) phi_j(x_q) JxW_q scattered into the stiffness matrix by `fdal_assemble_al_term`
(immersed_laplace.cc:659-702) against the same term formed with scipy on the host."""
import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import FdalError
from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import synthetic as syn
from fictitious_domain_al_preconditioners_b200.context import assemble_al_term

from . import parity_log as PL

pytestmark = pytest.mark.gpu


def _points(dim, nel, p, nq, seed):
    """Random 'immersed quadrature points' with weights, and the (dofs, phi) arrays of their background cells."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(0.05, 0.95, (nq, dim))
    w = rng.uniform(0.1, 1.0, nq)
    Phi = syn.background_shape_matrix(pts, nel, 0.0, 1.0, p).tocsr()
    dpc = (p + 1) ** dim
    dofs = -np.ones((nq, dpc), dtype=np.int32)
    phi = np.zeros((nq, dpc))
    for q in range(nq):
        s, e = Phi.indptr[q], Phi.indptr[q + 1]
        dofs[q, : e - s] = Phi.indices[s:e]
        phi[q, : e - s] = Phi.data[s:e]
    return Phi, w, dofs, phi


@pytest.mark.parametrize("dim,nel,p", [(2, 32, 1), (2, 12, 2), (3, 8, 1)])
def test_al_term_scatter_matches_scipy(dim, nel, p):
    n1 = p * nel + 1
    K1, M1 = syn.fe1d(nel, 1.0 / nel, p, p, 1, 1), syn.fe1d(nel, 1.0 / nel, p, p)
    lap = None
    for k in range(dim):
        t = syn.kron_all([K1 if kk == k else M1 for kk in range(dim)][::-1])
        lap = t if lap is None else lap + t
    A = syn._csr(lap)
    assert A.shape[0] == n1**dim
    Phi, w, dofs, phi = _points(dim, nel, p, 4000, 3)
    ref = syn._csr(A + Phi.T @ sp.diags(w) @ Phi)
    # the FE pattern of A already holds every cell coupling: same pattern before and after
    assert np.array_equal(ref.indptr, A.indptr) and np.array_equal(ref.indices, A.indices)
    got = assemble_al_term(A, dofs, phi, w)
    err = float(np.abs(got.data - ref.data).max() / np.abs(ref.data).max())
    PL.check(f"AL-term scatter dim={dim} p={p}", err, 1e-13)
    # deal.II row layout (diagonal first, rest ascending): the scatter searches rows linearly
    rp, ci, v = A.indptr.copy(), A.indices.copy(), A.data.copy()
    for i in range(A.shape[0]):
        s, e = rp[i], rp[i + 1]
        k = s + int(np.nonzero(ci[s:e] == i)[0][0])
        ci[s:k + 1] = np.roll(ci[s:k + 1], 1)
        v[s:k + 1] = np.roll(v[s:k + 1], 1)
    A2 = sp.csr_matrix((v, ci, rp), shape=A.shape)
    got2 = assemble_al_term(A2, dofs, phi, w)
    assert abs(got2 - ref).max() < 1e-13 * np.abs(ref.data).max()


def test_al_term_scatter_reports_a_missing_pattern_entry():
    A = sp.identity(50, format="csr")
    dofs = np.array([[3, 7]], dtype=np.int32)  # (3,7) is not in the pattern of the identity
    with pytest.raises(FdalError) as e:
        assemble_al_term(A, dofs, np.array([[0.5, 0.5]]), np.array([1.0]))
    assert e.value.status == b.ERR_SHAPE
