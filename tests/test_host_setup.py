"""Host-side setup stand-ins (synthetic deal.II-like blocks, smoothed-aggregation hierarchy):
structural checks that the solve path's inputs are what the reference would hand over."""
import numpy as np
import pytest
import scipy.sparse as sp

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import amg_setup as am
from fictitious_domain_al_preconditioners_b200 import synthetic as syn


def test_fe1d_matrices_are_the_textbook_ones():
    h = 0.25
    K = syn.fe1d(4, h, 1, 1, 1, 1).toarray()
    M = syn.fe1d(4, h, 1, 1).toarray()
    assert np.allclose(K[1, :3], np.array([-1, 2, -1]) / h)
    assert np.allclose(M[1, :3], np.array([1, 4, 1]) * h / 6)
    K2 = syn.fe1d(2, 0.5, 2, 2, 1, 1).toarray()
    assert np.allclose(K2[:3, :3], np.array([[7, -8, 1], [-8, 16, -8], [1, -8, 14]]) / (3 * 0.5))
    assert abs(syn.fe1d(3, 1 / 3, 2, 2).sum() - 1.0) < 1e-14  # mass sums to the length


@pytest.mark.parametrize("fac,kw", [(syn.immersed_laplace, dict(r_bg=4)),
                                    (syn.stokes_immersed_boundary, dict(dim=2, nel=8)),
                                    (syn.stokes_immersed_boundary, dict(dim=3, nel=4, r_emb=1)),
                                    (syn.elliptic_interface, dict(cycle=1)),
                                    (syn.elasticity_interface, dict(cycle=1))])
def test_blocks_have_the_structure_the_reference_assembles(fac, kw):
    p = fac(**kw)
    n, m = p.Ct.shape
    assert abs(p.A - p.A.T).max() < 1e-12 and abs(p.M - p.M.T).max() < 1e-14
    assert p.A.diagonal().min() > 0 and p.M.diagonal().min() > 0
    # DEBUG check of the reference: the coupling matrix integrates a partition of unity
    # (sum_i Ct[i, j] = int psi_j = (M 1)_j; nitsche_bcs.cc:467-490) wherever no Dirichlet row was cut
    assert np.allclose(np.asarray(p.Ct.sum(axis=0)).ravel(), p.M @ np.ones(m), rtol=1e-10, atol=1e-13)
    # constrained rows: only the diagonal (AffineConstraints::distribute_local_to_global)
    lone = np.diff(p.A.indptr) == 1
    assert lone.sum() > 0
    rows = np.nonzero(lone)[0]
    assert np.array_equal(p.A.indices[p.A.indptr[rows]], rows)
    if p.Bt is not None:
        assert p.Bt.shape == (n, p.Mp.shape[0])
        assert abs(p.Bt[rows]).sum() == 0  # B^T rows of constrained velocity DoFs are empty
        # div of a constant field is zero: B 1_c sums (over interior test functions) to ~0 against constant pressure
        assert abs((p.Mp @ np.ones(p.Mp.shape[0])).sum() - 1.0) < 1e-12  # |Omega| = 1
    if p.A2 is not None:
        assert abs(p.A2 - p.A2.T).max() < 1e-10
        assert np.abs(p.A2 @ np.ones(m)).max() < 1e-9 * abs(p.A2).max()  # pure Neumann block: constants in the kernel


def test_tensor_assembly_matches_kron_assembly():
    for dim, nel, num in ((2, 6, "component"), (3, 3, "node"), (2, 5, "node")):
        a = syn.stokes_immersed_boundary(dim=dim, nel=nel, r_emb=2 if dim == 2 else 1, fast=False, numbering=num)
        c = syn.stokes_immersed_boundary(dim=dim, nel=nel, r_emb=2 if dim == 2 else 1, fast=True, numbering=num)
        assert (a.A != c.A).nnz == 0
        assert abs(a.Bt - c.Bt).max() == 0 and abs(a.Ct - c.Ct).max() == 0 and np.array_equal(a.rhs, c.rhs)


def test_openmp_spgemm_matches_scipy():
    A = sp.random(40000, 2500, 0.025, random_state=2, format="csr")  # 2.5e6 non-zeros -> OpenMP path
    B = sp.random(2500, 1800, 0.01, random_state=3, format="csr")
    assert A.nnz >= 2_000_000
    C1 = am._spgemm(A, B)
    C2 = (A @ B).tocsr()
    assert abs(C1 - C2).max() < 1e-12
    assert np.all(np.diff(C1.indptr) >= 0) and C1.shape == C2.shape
    rows = np.repeat(np.arange(C1.shape[0]), np.diff(C1.indptr))
    assert np.all((np.diff(C1.indices) > 0) | (np.diff(rows) > 0))  # sorted, duplicate-free columns per row


def test_hierarchy_is_galerkin_and_coarsens():
    p = syn.stokes_immersed_boundary(dim=2, nel=16)
    H = syn.build_hierarchies(p, max_coarse=100)[b.AMG_A11]
    assert len(H.levels) >= 3
    for L, Ln in zip(H.levels[:-1], H.levels[1:]):
        assert abs(L.R - L.P.T).max() == 0
        assert abs(L.R @ L.A @ L.P - Ln.A).max() < 1e-10 * abs(Ln.A).max()
        assert Ln.A.shape[0] < L.A.shape[0] / 2
        lam = np.max(np.abs(np.linalg.eigvals((sp.diags(L.inv_diag) @ L.A).toarray()))) if L.A.shape[0] < 1500 else None
        if lam is not None:
            assert 0.7 * lam <= L.lambda_max <= 1.05 * lam
    # constant modes: aggregates never mix velocity components (utilities.h:304-309)
    comp = p.amg_comp[b.AMG_A11]
    P0 = H.levels[0].P.tocsc()
    S = am._strength_graph(H.levels[0].A, p.amg_theta[b.AMG_A11], comp)
    agg, n_agg = am.aggregate(S)
    for a_id in range(0, n_agg, max(1, n_agg // 50)):
        members = np.nonzero(agg == a_id)[0]
        assert len(set(comp[members])) == 1
    # Dirichlet rows have no strong connection: left out of the coarse grid
    assert np.all(agg[np.diff(H.levels[0].A.indptr) == 1] == -1)


def test_add_low_rank_rows_equals_the_general_sparse_add():
    """The explicit augmented matrix A + gamma Ct W^-1 C is formed by touching only the rows next to the immersed
    body (synthetic.add_low_rank_rows): bit-identical to scipy's general add + canonicalisation."""
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    A = sp.random(400, 400, density=0.03, random_state=7, format="csr") + sp.identity(400, format="csr")
    A = syn._csr(A)
    S = sp.lil_matrix((400, 400))
    for r in (0, 17, 18, 19, 250, 399):
        cols = rng.choice(400, 9, replace=False)
        S[r, cols] = rng.uniform(-1, 1, 9)
    S = sp.csr_matrix(S)
    ref = syn._csr(A + S)
    got = syn.add_low_rank_rows(A, S)
    assert np.array_equal(got.indptr, ref.indptr) and np.array_equal(got.indices, ref.indices)
    assert np.array_equal(got.data, ref.data)
    # nothing to add / an S touching most rows fall back to the general path
    assert (syn.add_low_rank_rows(A, sp.csr_matrix((400, 400))) != A).nnz == 0
    dense_S = sp.random(400, 400, density=0.02, random_state=3, format="csr")
    assert (syn.add_low_rank_rows(A, dense_S) != syn._csr(A + dense_S)).nnz == 0
