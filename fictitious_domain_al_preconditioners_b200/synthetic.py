"""Synthetic stand-ins for the deal.II side of the reference (host, setup only).

The reference meshes two non-matching grids, assembles the FE blocks and the
NonMatching coupling matrix with deal.II and then enters the solve path
(immersed_laplace.cc:278-496, stokes_immersed_boundary.cc:410-820,
elliptic_interface.cc:450-670).  deal.II is not installed here, so these
generators produce structurally faithful blocks on uniform grids — tensor
product Q1/Q2 assembly by Kronecker products, Dirichlet rows/columns eliminated
the way ``AffineConstraints::distribute_local_to_global`` leaves them (zeroed,
positive diagonal kept), coupling matrices by Gauss quadrature on the immersed
curve / surface / area with exact point location — for the named parameter files
(SURVEY.md 8(d)).  They feed tests and bench; nothing here is on the hot path.

Numbering: background scalar DoFs lexicographic (x fastest); vector-valued
background blocks component-wise ``[u_x | u_y | u_z]`` like
``DoFRenumbering::component_wise`` (stokes_immersed_boundary.cc:539-541);
immersed vector DoFs interleaved per vertex like ``FESystem``.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import _binding as b
from .context import ALConfig, ReductionControl, SolverControl

# ----------------------------------------------------------------------------- 1-D building blocks
_GAUSS = {n: np.polynomial.legendre.leggauss(n) for n in range(1, 9)}


def gauss01(n):
    x, w = _GAUSS[n]
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange(p: int, xi: np.ndarray, deriv: int = 0) -> np.ndarray:
    """Values (deriv=0) or first derivatives of the degree-p Lagrange basis on [0,1]
    with equispaced nodes; shape (len(xi), p+1)."""
    xi = np.asarray(xi, dtype=np.float64)
    nodes = np.linspace(0.0, 1.0, p + 1)
    out = np.zeros((xi.size, p + 1))
    for a in range(p + 1):
        others = [c for c in range(p + 1) if c != a]
        den = np.prod([nodes[a] - nodes[c] for c in others])
        if deriv == 0:
            num = np.ones_like(xi)
            for c in others:
                num = num * (xi - nodes[c])
            out[:, a] = num / den
        else:
            s = np.zeros_like(xi)
            for skip in others:
                t = np.ones_like(xi)
                for c in others:
                    if c != skip:
                        t = t * (xi - nodes[c])
                s += t
            out[:, a] = s / den
    return out


def fe1d(nel: int, h: float, p_test: int, p_trial: int, d_test: int = 0, d_trial: int = 0) -> sp.csr_matrix:
    """1-D FE matrix  int D^{d_test} phi_a D^{d_trial} psi_b  on a uniform grid."""
    xq, wq = gauss01(4)
    Te = lagrange(p_test, xq, d_test) * h ** (-d_test)
    Tr = lagrange(p_trial, xq, d_trial) * h ** (-d_trial)
    Me = (Te * (wq * h)[:, None]).T @ Tr
    e = np.arange(nel)
    rows = (e[:, None, None] * p_test + np.arange(p_test + 1)[None, :, None]) + np.zeros((1, 1, p_trial + 1), int)
    cols = (e[:, None, None] * p_trial + np.arange(p_trial + 1)[None, None, :]) + np.zeros((1, p_test + 1, 1), int)
    vals = np.broadcast_to(Me, rows.shape)
    A = sp.coo_matrix((vals.ravel(), (rows.ravel(), cols.ravel())), shape=(p_test * nel + 1, p_trial * nel + 1))
    A = A.tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def kron_all(mats):
    """kron(m[0], kron(m[1], ...)) — mats ordered slowest dimension first (z, y, x)."""
    out = mats[0]
    for m in mats[1:]:
        out = sp.kron(out, m, format="csr")
    return out.tocsr()


def boundary_mask(n1: int, dim: int) -> np.ndarray:
    """True on boundary nodes of an n1^dim lexicographic grid."""
    idx = np.indices((n1,) * dim)
    m = np.zeros((n1,) * dim, dtype=bool)
    for k in range(dim):
        m |= (idx[k] == 0) | (idx[k] == n1 - 1)
    return m.ravel()


def apply_dirichlet(A: sp.csr_matrix, constrained: np.ndarray) -> sp.csr_matrix:
    """Zero constrained rows and columns, keep the (positive) diagonal."""
    A = A.tocsr()
    d = A.diagonal()
    row = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    keep = ~(constrained[row] | constrained[A.indices])
    A = sp.csr_matrix((A.data[keep], A.indices[keep], np.concatenate([[0], np.cumsum(np.bincount(row[keep], minlength=A.shape[0]))])), shape=A.shape)
    A = A + sp.diags(np.where(constrained, d, 0.0))
    A = A.tocsr()
    A.sort_indices()
    return A


def zero_rows(A: sp.csr_matrix, constrained: np.ndarray) -> sp.csr_matrix:
    A = sp.diags((~constrained).astype(np.float64)) @ A
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    return A


# ----------------------------------------------------------------------------- point evaluation
def background_shape_matrix(points: np.ndarray, nel: int, lo: float, hi: float, p: int) -> sp.csr_matrix:
    """Phi[q, i] = phi_i(x_q) for the tensor-product degree-p space on [lo,hi]^dim."""
    nq, dim = points.shape
    h = (hi - lo) / nel
    n1 = p * nel + 1
    t = (points - lo) / h
    cell = np.clip(np.floor(t).astype(np.int64), 0, nel - 1)
    xi = t - cell
    vals = np.ones((nq, 1))
    idx = np.zeros((nq, 1), dtype=np.int64)
    # slowest dimension first so that x is fastest in the flattened index
    for k in range(dim - 1, -1, -1):
        s = lagrange(p, xi[:, k])  # (nq, p+1)
        node = cell[:, k][:, None] * p + np.arange(p + 1)[None, :]
        vals = (vals[:, :, None] * s[:, None, :]).reshape(nq, -1)
        idx = (idx[:, :, None] * n1 + node[:, None, :]).reshape(nq, -1)
    rows = np.repeat(np.arange(nq), vals.shape[1])
    Phi = sp.csr_matrix((vals.ravel(), (rows, idx.ravel())), shape=(nq, n1**dim))
    return Phi


# ----------------------------------------------------------------------------- immersed meshes
def circle_polyline(nseg: int, R: float, center):
    """Open parametrisation t in [0,1] of a circle: nseg segments, nseg+1 vertices
    (first and last coincide geometrically but are distinct DoFs, like the
    reference's mapped hyper_cube<1,2>, immersed_laplace.cc:311-323)."""
    t = np.arange(nseg + 1) / nseg
    X = np.stack([R * np.cos(2 * np.pi * t) + center[0], R * np.sin(2 * np.pi * t) + center[1]], axis=1)
    cells = np.stack([np.arange(nseg), np.arange(1, nseg + 1)], axis=1)
    return X, cells


def cubed_sphere(k: int, R: float, center):
    """Quadrilateral surface mesh of a sphere: 6 * 4^k cells, 6 * 4^k + 2 vertices
    (GridGenerator::hyper_sphere refined k times, stokes_immersed_boundary.cc:427)."""
    n = 2**k
    s = np.linspace(-1.0, 1.0, n + 1)
    u, v = np.meshgrid(s, s, indexing="ij")
    faces = []
    one = np.ones_like(u)
    for axis in range(3):
        for sign in (-1.0, 1.0):
            c = [None, None, None]
            c[axis] = sign * one
            c[(axis + 1) % 3] = u if sign > 0 else v
            c[(axis + 2) % 3] = v if sign > 0 else u
            faces.append(np.stack(c, axis=-1).reshape(-1, 3))
    pts = np.concatenate(faces)
    key = np.round(pts * n).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    inv = inv.ravel()
    verts = pts[first]
    verts = verts / np.linalg.norm(verts, axis=1)[:, None] * R + np.asarray(center)[None, :]
    cells = []
    for f in range(6):
        base = f * (n + 1) ** 2
        i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        a = base + i * (n + 1) + j
        cells.append(np.stack([a, a + (n + 1), a + (n + 1) + 1, a + 1], axis=-1).reshape(-1, 4))
    cells = inv[np.concatenate(cells)]
    return verts, cells


def disk_mesh(k: int, R: float, center=(0.0, 0.0)):
    """GridGenerator::hyper_ball (5 cells) refined k times; only boundary edges are
    curved (third generator argument ``false`` in parameters_ideal.prm:62)."""
    a = R / (1.0 + np.sqrt(2.0))
    d = R / np.sqrt(2.0)
    V = np.array([[-d, -d], [d, -d], [-a, -a], [a, -a], [-a, a], [a, a], [-d, d], [d, d]], dtype=np.float64)
    cells = np.array([[0, 1, 3, 2], [0, 2, 4, 6], [2, 3, 5, 4], [1, 7, 5, 3], [6, 4, 5, 7]], dtype=np.int64)
    bnd = {(0, 1), (1, 7), (6, 7), (0, 6)}
    bnd_edges = np.array(sorted(bnd), dtype=np.int64)
    for _ in range(k):
        nv = V.shape[0]
        e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [2, 3]], cells[:, [3, 0]]])
        es = np.sort(e, axis=1)
        ue, inv = np.unique(es, axis=0, return_inverse=True)
        inv = inv.ravel()
        mid = 0.5 * (V[ue[:, 0]] + V[ue[:, 1]])
        # boundary edges: project midpoint on the circle
        bkey = set(map(tuple, bnd_edges.tolist()))
        is_b = np.array([tuple(x) in bkey for x in ue.tolist()])
        nrm = np.linalg.norm(mid[is_b], axis=1)
        mid[is_b] *= (R / nrm)[:, None]
        nc = cells.shape[0]
        em = nv + inv.reshape(4, nc).T  # (nc, 4): midpoints of edges 01, 12, 23, 30
        cen = 0.5 * (V_ext(V, mid, em).sum(axis=1)) - 0.25 * V[cells].sum(axis=1)
        cid = nv + ue.shape[0] + np.arange(nc)
        V = np.concatenate([V, mid, cen])
        c0, c1, c2, c3 = cells.T
        m01, m12, m23, m30 = em.T
        cells = np.concatenate(
            [
                np.stack([c0, m01, cid, m30], 1),
                np.stack([m01, c1, m12, cid], 1),
                np.stack([cid, m12, c2, m23], 1),
                np.stack([m30, cid, m23, c3], 1),
            ]
        )
        # children of boundary edges stay boundary edges
        bidx = np.nonzero(is_b)[0]
        bnd_edges = np.sort(
            np.concatenate(
                [np.stack([ue[bidx, 0], nv + bidx], 1), np.stack([ue[bidx, 1], nv + bidx], 1)]
            ),
            axis=1,
        )
    return V + np.asarray(center)[None, :], cells


def V_ext(V, mid, em):
    allv = np.concatenate([V, mid])
    return allv[em]


def q1_quad_points(V, cells, nq):
    """Quadrature points, weights (JxW) and shape matrix of a bilinear quad mesh
    embedded in R^d (d = 2: area, d = 3: surface)."""
    xq, wq = gauss01(nq)
    xi, eta = np.meshgrid(xq, xq, indexing="ij")
    w2 = np.outer(wq, wq).ravel()
    xi, eta = xi.ravel(), eta.ravel()
    N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), xi * eta, (1 - xi) * eta], axis=1)  # (q,4)
    dNx = np.stack([-(1 - eta), (1 - eta), eta, -eta], axis=1)
    dNe = np.stack([-(1 - xi), -xi, xi, (1 - xi)], axis=1)
    X = V[cells]  # (nc,4,d)
    pts = np.einsum("qa,cad->cqd", N, X)
    t1 = np.einsum("qa,cad->cqd", dNx, X)
    t2 = np.einsum("qa,cad->cqd", dNe, X)
    d = V.shape[1]
    if d == 2:
        J = t1[..., 0] * t2[..., 1] - t1[..., 1] * t2[..., 0]
    else:
        J = np.linalg.norm(np.cross(t1, t2), axis=-1)
    JxW = np.abs(J) * w2[None, :]
    nc = cells.shape[0]
    rows = np.repeat(np.arange(nc * N.shape[0]), 4)
    cols = np.repeat(cells, N.shape[0], axis=0).reshape(nc, N.shape[0], 4).ravel()
    vals = np.broadcast_to(N[None], (nc,) + N.shape).ravel()
    Psi = sp.csr_matrix((vals, (rows, cols)), shape=(nc * N.shape[0], V.shape[0]))
    extra = dict(N=N, dNx=dNx, dNe=dNe, t1=t1, t2=t2, J=J)
    return pts.reshape(-1, d), JxW.ravel(), Psi, extra


def segment_points(V, cells, nq):
    xq, wq = gauss01(nq)
    X0, X1 = V[cells[:, 0]], V[cells[:, 1]]
    length = np.linalg.norm(X1 - X0, axis=1)
    pts = X0[:, None, :] * (1 - xq)[None, :, None] + X1[:, None, :] * xq[None, :, None]
    JxW = length[:, None] * wq[None, :]
    nc = cells.shape[0]
    rows = np.repeat(np.arange(nc * nq), 2)
    cols = np.stack([np.repeat(cells[:, 0], nq), np.repeat(cells[:, 1], nq)], axis=1).ravel()
    vals = np.stack([np.tile(1 - xq, nc), np.tile(xq, nc)], axis=1).ravel()
    Psi = sp.csr_matrix((vals, (rows, cols)), shape=(nc * nq, V.shape[0]))
    return pts.reshape(-1, V.shape[1]), JxW.ravel(), Psi, length


def q1_quad_stiffness(V, cells, nq=2):
    """(grad psi_a, grad psi_b) on a planar bilinear quad mesh."""
    pts, JxW, Psi, ex = q1_quad_points(V, cells, nq)
    nc = cells.shape[0]
    t1, t2, J = ex["t1"], ex["t2"], ex["J"]
    # inverse Jacobian: [dxi/dx dxi/dy; deta/dx deta/dy]
    inv = np.empty(t1.shape[:2] + (2, 2))
    inv[..., 0, 0] = t2[..., 1] / J
    inv[..., 0, 1] = -t2[..., 0] / J
    inv[..., 1, 0] = -t1[..., 1] / J
    inv[..., 1, 1] = t1[..., 0] / J
    gx = ex["dNx"][None, :, :, None] * inv[:, :, None, 0, :] + ex["dNe"][None, :, :, None] * inv[:, :, None, 1, :]
    Ke = np.einsum("cqad,cqbd,cq->cab", gx, gx, JxW.reshape(nc, -1))
    rows = np.repeat(cells[:, :, None], 4, axis=2).ravel()
    cols = np.repeat(cells[:, None, :], 4, axis=1).ravel()
    K = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(V.shape[0], V.shape[0])).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def _csr(A):
    A = sp.csr_matrix(A)
    A.sum_duplicates()
    A.sort_indices()
    return A


def add_low_rank_rows(A: sp.csr_matrix, S: sp.csr_matrix) -> sp.csr_matrix:
    """A + S for a huge A and an S that only touches a few rows (the AL term gamma Ct W^-1 C lives on the
    rows next to the immersed body): the untouched rows of A are copied as they are, only the touched rows
    go through a sparse add — no sort / duplicate pass over the 10^9 entries of A (the general scipy
    `A + S` followed by canonicalisation costs tens of seconds there)."""
    A = A.tocsr()
    S = sp.csr_matrix(S)
    S.sum_duplicates()
    S.sort_indices()
    rows = np.nonzero(np.diff(S.indptr))[0]
    if rows.size == 0 or rows.size > A.shape[0] // 2 or not A.has_sorted_indices:
        return _csr(A + S)
    sub = (A[rows] + S[rows]).tocsr()
    sub.sort_indices()
    cnt = np.diff(A.indptr).astype(np.int64)
    cnt[rows] = np.diff(sub.indptr)
    indptr = np.zeros(A.shape[0] + 1, dtype=np.int64)
    np.cumsum(cnt, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.empty(nnz, dtype=A.indices.dtype)
    data = np.empty(nnz, dtype=np.float64)
    # copy the untouched stretches between consecutive touched rows in bulk
    r_prev = 0
    for k, r in enumerate(np.append(rows, A.shape[0])):
        if r > r_prev:  # rows [r_prev, r) unchanged
            a0, a1 = int(A.indptr[r_prev]), int(A.indptr[r])
            o0 = int(indptr[r_prev])
            indices[o0:o0 + (a1 - a0)] = A.indices[a0:a1]
            data[o0:o0 + (a1 - a0)] = A.data[a0:a1]
        if r < A.shape[0]:
            s0, s1 = int(sub.indptr[k]), int(sub.indptr[k + 1])
            o0 = int(indptr[r])
            indices[o0:o0 + (s1 - s0)] = sub.indices[s0:s1]
            data[o0:o0 + (s1 - s0)] = sub.data[s0:s1]
        r_prev = r + 1
    out = sp.csr_matrix((data, indices, indptr if nnz >= 2**31 - 1 else indptr.astype(np.int32)), shape=A.shape)
    out.has_sorted_indices = True
    out.has_canonical_format = True
    return out


# ----------------------------------------------------------------------------- problem container
@dataclass
class Problem:
    name: str
    config: ALConfig
    A: sp.csr_matrix
    Ct: sp.csr_matrix
    M: sp.csr_matrix
    rhs: np.ndarray  # un-augmented block right-hand side
    A2: sp.csr_matrix | None = None
    Bt: sp.csr_matrix | None = None
    Mp: sp.csr_matrix | None = None
    winv_diag: np.ndarray | None = None
    amg_matrix: dict = field(default_factory=dict)  # which -> explicit matrix AMG is built on
    amg_theta: dict = field(default_factory=dict)
    amg_comp: dict = field(default_factory=dict)
    augment_rhs: bool = True
    meta: dict = field(default_factory=dict)

    @property
    def sizes(self):
        k = self.config.kind
        n, m = self.Ct.shape
        if k == b.KIND_LAPLACE:
            return (n, m)
        if k in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES):
            return (n, self.Bt.shape[1], m)
        return (n, m, m)

    @property
    def n_dofs(self):
        return int(sum(self.sizes))


def _winv_diag_from(M, squared):
    d = M.diagonal()
    return 1.0 / (d * d) if squared else 1.0 / d


# ----------------------------------------------------------------------------- C1: immersed_laplace
def immersed_laplace(
    r_bg: int = 6,
    r_emb: int | None = None,
    diagonal_inverse: bool = True,
    operator_form: bool = False,
    f: float = 1.0,
    g: float = 1.0,
    R: float = 0.2,
    center=(0.4, 0.4),
    nq_coupling: int = 3,
) -> Problem:
    """2x2 system of immersed_laplace with ``Solver = augmented``
    (parameters/circle/*.prm; immersed_laplace.cc:636-948)."""
    if r_emb is None:
        # the reference refines the background once more near the curve
        # (delta_refinement = 1) and the embedded grid one level above the global
        # background level; on a uniform surrogate that is h_emb ~ 1.26 h_bg
        r_emb = r_bg
    nel = 2**r_bg
    h = 1.0 / nel
    K1, M1 = fe1d(nel, h, 1, 1, 1, 1), fe1d(nel, h, 1, 1)
    A = kron_all([M1, K1]) + kron_all([K1, M1])
    Mbg = kron_all([M1, M1])
    bnd = boundary_mask(nel + 1, 2)
    A = apply_dirichlet(_csr(A), bnd)
    Xc, cells = circle_polyline(2**r_emb, R, center)
    pts, JxW, Psi, length = segment_points(Xc, cells, nq_coupling)
    Phi = background_shape_matrix(pts, nel, 0.0, 1.0, 1)
    W = sp.diags(JxW)
    Ct = zero_rows(_csr(Phi.T @ W @ Psi), bnd)
    M = _csr(Psi.T @ W @ Psi)
    m = M.shape[0]
    n = A.shape[0]
    fvec = f * (Mbg @ np.ones(n))
    fvec[bnd] = 0.0
    gvec = g * (M @ np.ones(m))
    gamma = 10.0  # immersed_laplace.cc:647
    cfg = ALConfig(
        kind=b.KIND_LAPLACE,
        restart=30,
        gamma=gamma,
        inner=SolverControl(100, 1e-2),  # :907
        outer=ReductionControl(1000, 1e-10, 1e-12),  # circle prm + ctor default
    )
    prob = Problem(name=f"immersed_laplace_r{r_bg}", config=cfg, A=A, Ct=Ct, M=M, rhs=np.concatenate([fvec, gvec]))
    if operator_form:
        # immersed_laplace.cc:653-705: gamma /= h_immersed, A += gamma * sum phi_i phi_j JxW
        h_imm = float(length.max())
        cfg.gamma = gamma / h_imm
        Dint = sp.diags((~bnd).astype(np.float64))
        G = _csr(Dint @ (Phi.T @ W @ Phi) @ Dint)
        G.eliminate_zeros()
        prob.A = _csr(A + cfg.gamma * G)
        cfg.aug_explicit = True
        if diagonal_inverse:
            cfg.winv_mode = b.WINV_DIAG
            prob.winv_diag = _winv_diag_from(M, squared=False)
        else:
            cfg.winv_mode = b.WINV_EXACT_M
        prob.amg_matrix[b.AMG_A11] = prob.A  # :704
    else:
        if diagonal_inverse:
            cfg.winv_mode = b.WINV_DIAG
            prob.winv_diag = _winv_diag_from(M, squared=True)  # :869-873
        else:
            cfg.winv_mode = b.WINV_EXACT_M_SQUARED  # :875-876
        # AMG on A + gamma Ct diag(1/Mii^2) C in both cases (:712-715, 815-833)
        D = sp.diags(_winv_diag_from(M, squared=True))
        prob.amg_matrix[b.AMG_A11] = _csr(A + cfg.gamma * (Ct @ D @ Ct.T))
    prob.amg_theta[b.AMG_A11] = 1e-4
    prob.meta = dict(h=h, n_bg=n, m=m, r_bg=r_bg, r_emb=r_emb)
    return prob


def nitsche_bcs(r: int = 5, multiplier_degree: int = 1, gamma: float = 10.0, manufactured: bool = True) -> Problem:
    """2x2 system of the fourth application, nitsche_bcs (Dirichlet data imposed weakly by a
    boundary multiplier; "next" row N4 of SURVEY 8(f)): the SAME BlockPreconditionerAugmented-
    Lagrangian path as immersed_laplace in operator form (nitsche_bcs.cc:497-661).

    Unit square, Q1 bulk space WITHOUT strong boundary conditions, multiplier space on the boundary
    mesh extracted from the bulk mesh (matching): continuous P1 (``multiplier_degree=1``) or the
    shipped discontinuous P0 (``=0``, parameters_nitsche.prm — its coupling matrix has the
    alternating multiplier in its kernel, so only the Krylov solve is meaningful there).
      stiffness  A = (grad u, grad v) + (gamma/h) <u, v>_boundary     (:531-566, explicit AL term)
      invW       = (1/h) M_b^-1  (UMFPACK)                             (:637-640)
      P          = BlockPreconditionerAugmentedLagrangian{A_inv, ., ., invW, gamma}  (:642)
    which is kind LAPLACE with aug_explicit, WINV_EXACT_M and gamma_eff = gamma / h.  The right-hand
    side carries the consistent augmentation (gamma/h) <g, v>_boundary itself (:575-632)."""
    nel = 2**r
    h = 1.0 / nel
    n1 = nel + 1
    K1, M1 = fe1d(nel, h, 1, 1, 1, 1), fe1d(nel, h, 1, 1)
    A0 = kron_all([M1, K1]) + kron_all([K1, M1])
    Mbg = kron_all([M1, M1])
    n = n1 * n1
    # boundary loop, counter-clockwise from the origin: 4 nel nodes / 4 nel edges
    k = np.arange(nel)
    ix = np.concatenate([k, np.full(nel, nel), nel - k, np.zeros(nel, int)])
    iy = np.concatenate([np.zeros(nel, int), k, np.full(nel, nel), nel - k])
    loop = ix + n1 * iy
    nb = 4 * nel
    E = sp.csr_matrix((np.ones(nb), (loop, np.arange(nb))), shape=(n, nb))  # bulk dof <- boundary node
    nxt = (np.arange(nb) + 1) % nb
    # P1 trace mass on the closed loop
    Mtrace = sp.coo_matrix(
        (np.concatenate([np.full(nb, 2 * h / 3), np.full(nb, h / 6), np.full(nb, h / 6)]),
         (np.concatenate([np.arange(nb), np.arange(nb), nxt]), np.concatenate([np.arange(nb), nxt, np.arange(nb)]))),
        shape=(nb, nb)).tocsr()
    if multiplier_degree == 1:
        M = _csr(Mtrace)
        Ct = _csr(E @ Mtrace)  # <phi_i, mu_j>: the trace of a bulk hat function is the boundary hat function
    else:
        M = _csr(sp.diags(np.full(nb, h)))
        # edge j joins boundary nodes j and j+1: int_e phi = h/2 for both
        Half = sp.coo_matrix((np.full(2 * nb, h / 2), (np.concatenate([np.arange(nb), nxt]), np.tile(np.arange(nb), 2))),
                             shape=(nb, nb)).tocsr()
        Ct = _csr(E @ Half)
    g_eff = gamma / h  # invW_scale = 1 / h_immersed (:519-522)
    A = _csr(A0 + g_eff * (E @ Mtrace @ E.T))
    X, Y = np.meshgrid(np.linspace(0, 1, n1), np.linspace(0, 1, n1), indexing="xy")
    x, y = X.ravel(), Y.ravel()
    if manufactured:
        u_exact = np.sin(np.pi * x) * np.cos(np.pi * y) + x
        f_nodal = 2 * np.pi**2 * np.sin(np.pi * x) * np.cos(np.pi * y)
    else:  # parameters_nitsche.prm: f = 1, g = x^2 + y^2
        u_exact = None
        f_nodal = np.ones(n)
    g_nodal = (u_exact if manufactured else x * x + y * y)[loop]
    rhs0 = Mbg @ f_nodal + g_eff * (E @ (Mtrace @ g_nodal))
    rhs1 = Ct.T @ (E @ g_nodal)  # <g_h, mu_j> with g_h the nodal interpolant on the boundary
    cfg = ALConfig(
        kind=b.KIND_LAPLACE,
        restart=30,
        gamma=g_eff,
        aug_explicit=True,
        winv_mode=b.WINV_EXACT_M,
        inner=ReductionControl(1000, 1e-2, 1e-10),   # parameters_nitsche.prm "Inner solver control"
        outer=ReductionControl(1000, 1e-12, 1e-9),   # "Outer solver control"
    )
    prob = Problem(name=f"nitsche_bcs_r{r}_p{multiplier_degree}", config=cfg, A=A, Ct=Ct, M=M,
                   rhs=np.concatenate([rhs0, rhs1]), augment_rhs=False)
    prob.amg_matrix[b.AMG_A11] = prob.A  # amg_prec.initialize(stiffness_matrix) (:568)
    prob.amg_theta[b.AMG_A11] = 1e-4
    prob.meta = dict(h=h, n_bg=n, m=nb, u_exact=u_exact, boundary_dofs=loop)
    return prob


# ----------------------------------------------------------------------------- fast velocity block
def velocity_block_tensor(nel: int, dim: int, gamma_grad_div: float, bnd_s: np.ndarray, device=None,
                          interleaved: bool = False, p: int = 2, length: float = 1.0, terms=None) -> sp.csr_matrix:
    """(grad u, grad v) + gamma_gd (div u, div v) on the Q2^dim tensor grid with the Dirichlet
    rows / columns already eliminated — the same matrix ``apply_dirichlet(bmat(kron ...))``
    builds, assembled by index arithmetic on torch tensors (on the GPU when there is one:
    at 10^9 non-zeros scipy's kron/bmat needs minutes and several copies of the matrix).
    All d*d blocks share one sparsity pattern (the tensor product of the 1-D Q2 pattern), so
    the pattern is sorted into CSR order once and every block is a product of 1-D values."""
    import torch

    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    h = length / nel
    n1 = p * nel + 1
    ns = n1**dim
    K2, M2, D10 = (fe1d(nel, h, p, p, 1, 1), fe1d(nel, h, p, p), fe1d(nel, h, p, p, 1, 0))
    if terms is None:
        # Stokes velocity block: (grad u, grad v) + gamma_gd (div u, div v)
        def terms(c, d):
            if c == d:
                t = [(1.0, ["K" if kk == k else "M" for kk in range(dim)]) for k in range(dim)]
                return t + [(gamma_grad_div, ["K" if kk == c else "M" for kk in range(dim)])]
            return [(gamma_grad_div, ["D" if kk == c else ("Dt" if kk == d else "M") for kk in range(dim)])]
    pat = (abs(K2) + abs(M2) + abs(D10) + abs(D10.T)).tocoo()
    order = np.lexsort((pat.col, pat.row))
    r1 = torch.from_numpy(pat.row[order].astype(np.int64)).to(device)
    c1 = torch.from_numpy(pat.col[order].astype(np.int64)).to(device)
    dense = {"K": K2.toarray(), "M": M2.toarray(), "D": D10.toarray(), "Dt": D10.T.toarray()}
    v1 = {k: torch.from_numpy(np.ascontiguousarray(a[pat.row[order], pat.col[order]])).to(device) for k, a in dense.items()}
    nn = r1.numel()

    def tensor_index(t1):  # flattened tensor-product index, slowest dimension first
        out = t1
        for _ in range(dim - 1):
            out = (out[:, None] * n1 + t1[None, :]).reshape(-1)
        return out

    row3 = tensor_index(r1)
    col3 = tensor_index(c1)
    perm = torch.argsort(row3 * ns + col3)
    row3 = row3[perm]
    col3 = col3[perm]
    bnd = torch.from_numpy(bnd_s).to(device)
    interior = ~(bnd[row3] | bnd[col3])
    diag_keep = interior | (row3 == col3)

    def block_vals(names):  # names[k]: factor of dimension k (0 = x, fastest)
        out = v1[names[dim - 1]]
        for k in range(dim - 2, -1, -1):
            out = (out[:, None] * v1[names[k]][None, :]).reshape(-1)
        return out[perm]

    def block(c, d):
        out = None
        for coef, names in terms(c, d):
            t = coef * block_vals(names)
            out = t if out is None else out + t
        return out

    if interleaved:
        # node-major numbering (row = node*dim + c): all d*d blocks of an interior row have
        # the same pattern, so entry j of block (c,d) sits at indptr[row] + j*dim + d; a
        # boundary row keeps only its diagonal
        rows_i, cols_i = row3[interior], col3[interior]
        cnt_int = torch.bincount(rows_i, minlength=ns)
        blkptr = torch.zeros(ns + 1, dtype=torch.int64, device=device)
        blkptr[1:] = torch.cumsum(cnt_int, 0)
        j = torch.arange(rows_i.numel(), device=device) - blkptr[rows_i]
        row_cnt = torch.where(bnd, torch.ones_like(cnt_int), cnt_int * dim)
        row_cnt = row_cnt.repeat_interleave(dim)
        n = dim * ns
        indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
        indptr[1:] = torch.cumsum(row_cnt, 0)
        nnz = int(indptr[-1].item())
        indices = torch.empty(nnz, dtype=torch.int32, device=device)
        data = torch.empty(nnz, dtype=torch.float64, device=device)
        bd = (row3 == col3) & bnd[row3]
        rows_b = row3[bd]
        for c in range(dim):
            base = indptr[rows_i * dim + c] + j * dim
            for d in range(dim):
                v = block(c, d)
                if c == d:
                    pb = indptr[rows_b * dim + c]
                    indices[pb] = (rows_b * dim + c).to(torch.int32)
                    data[pb] = v[bd]
                indices[base + d] = (cols_i * dim + d).to(torch.int32)
                data[base + d] = v[interior]
                del v
            del base
        A = sp.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(),
                           indptr.cpu().numpy().astype(np.int64 if nnz >= 2**31 - 1 else np.int32)), shape=(n, n))
        A.has_sorted_indices = True
        return A
    # per component row block: entries of blocks d = 0..dim-1, merged into CSR order
    cnt = []
    rows_k, cols_k, vals_k = [], [], []
    for c in range(dim):
        for d in range(dim):
            v = block(c, d)
            keep = diag_keep if c == d else interior
            rows_k.append(row3[keep])
            cols_k.append(col3[keep])
            vals_k.append(v[keep])
            cnt.append(torch.bincount(rows_k[-1], minlength=ns))
            del v
    n = dim * ns
    row_cnt = torch.cat([sum(cnt[c * dim + d] for d in range(dim)) for c in range(dim)])
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(row_cnt, 0)
    nnz = int(indptr[-1].item())
    indices = torch.empty(nnz, dtype=torch.int32, device=device)
    data = torch.empty(nnz, dtype=torch.float64, device=device)
    for c in range(dim):
        off = torch.zeros(ns, dtype=torch.int64, device=device)
        for d in range(dim):
            k = c * dim + d
            blkptr = torch.zeros(ns + 1, dtype=torch.int64, device=device)
            blkptr[1:] = torch.cumsum(cnt[k], 0)
            rk = rows_k[k]
            pos = indptr[c * ns + rk] + off[rk] + (torch.arange(rk.numel(), device=device) - blkptr[rk])
            indices[pos] = (cols_k[k] + d * ns).to(torch.int32)
            data[pos] = vals_k[k]
            off = off + cnt[k]
            rows_k[k] = cols_k[k] = vals_k[k] = None
            del pos, rk, blkptr
    A = sp.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy().astype(np.int64 if nnz >= 2**31 - 1 else np.int32)),
                      shape=(n, n))
    A.has_sorted_indices = True
    return A


# ----------------------------------------------------------------------------- C2 / C4: Stokes
def stokes_immersed_boundary(
    dim: int = 2,
    r_bg: int | None = None,
    nel: int | None = None,
    r_emb: int | None = None,
    gamma: float = 10.0,
    gamma_grad_div: float = 10.0,
    diagonal_mass: bool | None = None,
    diag_minres: bool = False,
    nq_coupling: int | None = None,
    build_amg_matrix: bool = True,
    fast: bool | None = None,
    numbering: str = "component",
    grad_div_stabilization: bool = True,
) -> Problem:
    """3x3 system of stokes_immersed_boundary with ``Solver = IBStokesAL``
    (parameters_stokes.prm in 2-D, parameters_stokes_3d.prm in 3-D;
    stokes_immersed_boundary.cc:918-1079).  Q2^dim-Q1 Taylor-Hood on [0,1]^dim."""
    if nel is None:
        nel = 2 ** (r_bg if r_bg is not None else 4)
    h = 1.0 / nel
    if diagonal_mass is None:
        diagonal_mass = dim == 3  # parameters_stokes.prm:21 false, parameters_stokes_3d.prm:18 true
    gamma_gd_operator = gamma_grad_div
    if not grad_div_stabilization:
        # "Grad-div stabilization = false": the (div,div) term is not assembled into A; the operator
        # carries gamma_gd Bt Mp^-1 B instead and the inner CG runs unpreconditioned
        # (stokes_immersed_boundary.cc:992-995, 1046-1051)
        gamma_grad_div = 0.0
        build_amg_matrix = False
    K2, M2 = fe1d(nel, h, 2, 2, 1, 1), fe1d(nel, h, 2, 2)
    D10 = fe1d(nel, h, 2, 2, 1, 0)
    F = fe1d(nel, h, 1, 2, 0, 0)
    E = fe1d(nel, h, 1, 2, 0, 1)
    Mq1 = fe1d(nel, h, 1, 1)
    n1 = 2 * nel + 1
    ns = n1**dim
    dims = list(range(dim))  # 0 = x ... ; kron order slowest first
    order = dims[::-1]

    def K(facs):
        return kron_all([facs[k] for k in order])

    bnd_s = boundary_mask(n1, dim)
    bnd = np.tile(bnd_s, dim)
    if fast is None:
        fast = ns * dim >= 200_000
    import time as _time

    _t0 = _time.perf_counter()
    node_major = numbering == "node"
    # new -> old order of the node-major numbering (row = node*dim + c)
    node_order = (np.arange(ns, dtype=np.int64)[:, None] + ns * np.arange(dim, dtype=np.int64)[None, :]).reshape(-1)
    if fast:
        A = velocity_block_tensor(nel, dim, gamma_grad_div, bnd_s, interleaved=node_major)
        from .amg_setup import _log

        _log(f"velocity block nel={nel} dim={dim}: nnz={A.nnz} in {_time.perf_counter()-_t0:.2f}s")
    else:
        lap = None
        for k in dims:
            t = K({kk: (K2 if kk == k else M2) for kk in dims})
            lap = t if lap is None else lap + t
        blocks = [[None] * dim for _ in range(dim)]
        for c in dims:
            for d in dims:
                if c == d:
                    gd = K({kk: (K2 if kk == c else M2) for kk in dims})
                    blocks[c][d] = lap + gamma_grad_div * gd
                else:
                    gd = K({kk: (D10 if kk == c else (D10.T.tocsr() if kk == d else M2)) for kk in dims})
                    blocks[c][d] = gamma_grad_div * gd
        A = _csr(sp.bmat(blocks, format="csr"))
        del blocks, lap
        A = apply_dirichlet(A, bnd)
        if node_major:
            A = _csr(A[node_order][:, node_order])
    from .amg_setup import _log as _slog

    _t1 = _time.perf_counter()
    Bblk = [-K({kk: (E if kk == c else F) for kk in dims}) for c in dims]
    B = _csr(sp.hstack(Bblk, format="csr"))
    Mp = _csr(kron_all([Mq1] * dim))
    Bt = zero_rows(_csr(B.T), bnd)
    if node_major:
        Bt = _csr(Bt[node_order])
    n_u, n_p = A.shape[0], Mp.shape[0]
    _slog(f"B, Bt, Mp: {_time.perf_counter()-_t1:.2f}s")
    _t1 = _time.perf_counter()
    # immersed boundary
    if dim == 2:
        if r_emb is None:
            r_emb = int(round(np.log2(nel)))  # h_emb ~ 1.3 h_bg, see immersed_laplace()
        Xc, cells = circle_polyline(2**r_emb, 0.21, (0.45, 0.45))  # parameters_stokes.prm
        nq = nq_coupling or 3
        pts, JxW, Psi, _ = segment_points(Xc, cells, nq)
        gval = np.array([-0.5, 0.5])
        fval = np.array([1.0, 1.0])
    else:
        if r_emb is None:
            r_emb = max(1, int(round(np.log2(nel))) - 3)  # nel=128 -> 4 (parameters_stokes_3d.prm:8-10)
        Xc, cells = cubed_sphere(r_emb, 0.1, (0.5, 0.5, 0.5))  # stokes_immersed_boundary.cc:427
        nq = nq_coupling or 3
        pts, JxW, Psi, _ = q1_quad_points(Xc, cells, nq)
        gval = np.array([-0.5, 0.5, 0.0])
        fval = np.array([1.0, 0.0, 0.0])  # parameters_stokes_3d.prm body force
    Phi = background_shape_matrix(pts, nel, 0.0, 1.0, 2)
    W = sp.diags(JxW)
    Cs = zero_rows(_csr(Phi.T @ W @ Psi), bnd_s)  # scalar coupling (ns x ms)
    Ms = _csr(Psi.T @ W @ Psi)
    ms = Ms.shape[0]
    # vector-valued: background component-wise, immersed interleaved (j*dim + c)
    sel = [sp.csr_matrix((np.ones(ms), (np.arange(ms), np.arange(ms) * dim + c)), shape=(ms, ms * dim)) for c in dims]
    Ct = _csr(sp.vstack([Cs @ sel[c] for c in dims], format="csr"))
    if node_major:
        Ct = _csr(Ct[node_order])
    M = _csr(sum(sel[c].T @ Ms @ sel[c] for c in dims))
    m = M.shape[0]
    mass_u = kron_all([M2] * dim) @ np.ones(ns)
    fvec = np.concatenate([fval[c] * np.where(bnd_s, 0.0, mass_u) for c in dims])
    if node_major:
        fvec = fvec[node_order]
    gnod = np.tile(gval, ms)
    gvec = M @ gnod
    cfg = ALConfig(
        kind=b.KIND_STOKES_DIAG_MINRES if diag_minres else b.KIND_STOKES,
        restart=30,
        gamma=gamma,
        gamma_grad_div=gamma_gd_operator,
        grad_div_in_operator=not grad_div_stabilization,
        inner_prec=b.PREC_AMG if grad_div_stabilization else b.PREC_IDENTITY,
        inner=SolverControl(100, 1e-2),  # ALControl: Max steps 100, tol_AL 1e-2
        outer=ReductionControl(1000, 1e-8, 1e-12),
        mass=SolverControl(100, 1e-6),  # stokes_immersed_boundary.cc:934
    )
    winv = _winv_diag_from(M, squared=True)  # :976-978
    prob = Problem(
        name=f"stokes_ib_{dim}d_nel{nel}",
        config=cfg,
        A=A,
        Ct=Ct,
        M=M,
        Bt=Bt,
        Mp=Mp,
        rhs=np.concatenate([fvec, np.zeros(n_p), gvec]),
    )
    if diagonal_mass:
        cfg.winv_mode = b.WINV_DIAG
        cfg.mp_inv_mode = b.MPINV_CG_LUMPED
        prob.winv_diag = winv
    else:
        cfg.winv_mode = b.WINV_EXACT_M_SQUARED
        cfg.mp_inv_mode = b.MPINV_EXACT
    _slog(f"coupling blocks, rhs: {_time.perf_counter()-_t1:.2f}s")
    _t1 = _time.perf_counter()
    if build_amg_matrix:
        # build_AMG_augmented_block: (grad,grad)+gamma(div,div) + gamma Ct diag(1/Mii^2) C
        # (utilities.h:112-331; the AL gamma is used for grad-div, SURVEY Q7 — equal here)
        prob.amg_matrix[b.AMG_A11] = add_low_rank_rows(A, gamma * (Ct @ sp.diags(winv) @ Ct.T))
        _slog(f"explicit augmented matrix: {_time.perf_counter()-_t1:.2f}s")
        prob.amg_theta[b.AMG_A11] = 0.02  # utilities.h:314
        comp = np.repeat(np.arange(dim, dtype=np.int32), ns)
        prob.amg_comp[b.AMG_A11] = comp[node_order] if node_major else comp
    prob.meta = dict(h=h, n_u=n_u, n_p=n_p, m=m, nel=nel, dim=dim, r_emb=r_emb, node_major=node_major,
                     block_size=dim)
    return prob


# ----------------------------------------------------------------------------- C3: elliptic_interface
def elliptic_interface(
    cycle: int = 2,
    beta1: float = 1.0,
    beta2: float = 1e3,
    gamma_fluid: float = 10.0,
    gamma_solid: float = 1e-2,
    modified: bool = True,
    diagonal_inverse: bool = False,
    h_scaled: bool = True,
    fixed_iterations: bool = False,
    nq_coupling: int = 3,
) -> Problem:
    """3x3 system of elliptic_interface (parameters_elliptic_interface/parameters_ideal.prm;
    elliptic_interface.cc:676-956): background Q1 on [-1,1]^2 at refinement 4+cycle,
    immersed disk R=0.3 (hyper_ball) at refinement ``cycle``."""
    from .context import IterationNumberControl

    nel = 2 ** (4 + cycle)
    h = 2.0 / nel
    K1, M1 = fe1d(nel, h, 1, 1, 1, 1), fe1d(nel, h, 1, 1)
    lap = _csr(kron_all([M1, K1]) + kron_all([K1, M1]))
    Mbg = kron_all([M1, M1])
    bnd = boundary_mask(nel + 1, 2)
    A1 = apply_dirichlet(_csr(beta1 * lap), bnd)
    V, cells = disk_mesh(cycle, 0.3)
    pts, JxW, Psi, _ = q1_quad_points(V, cells, nq_coupling)
    Phi = background_shape_matrix(pts, nel, -1.0, 1.0, 1)
    W = sp.diags(JxW)
    Ct = zero_rows(_csr(Phi.T @ W @ Psi), bnd)
    M = _csr(Psi.T @ W @ Psi)
    A2 = _csr((beta2 - beta1) * q1_quad_stiffness(V, cells, 2))
    n, m = Ct.shape
    # maximal_cell_diameter of the immersed grid
    diag1 = np.linalg.norm(V[cells[:, 0]] - V[cells[:, 2]], axis=1)
    diag2 = np.linalg.norm(V[cells[:, 1]] - V[cells[:, 3]], axis=1)
    h_imm = float(np.maximum(diag1, diag2).max())
    if h_scaled:
        g1, g2 = gamma_fluid / h_imm**2, gamma_solid / h_imm**2  # :744-748
    else:
        g1, g2 = gamma_fluid, gamma_solid
    f1 = Mbg @ np.ones(n)
    f1[bnd] = 0.0
    f2 = M @ np.ones(m)
    inner = IterationNumberControl(30, 1e-4) if fixed_iterations else ReductionControl(100000, 1e-2, 1e-20)
    cfg = ALConfig(
        kind=b.KIND_ELLIPTIC_MODIFIED if modified else b.KIND_ELLIPTIC_IDEAL,
        restart=50,  # :863
        gamma=g1,
        gamma2=g2,
        inner=inner,
        outer=ReductionControl(1000, 1e-10, 1e-10),
    )
    prob = Problem(
        name=f"elliptic_interface_c{cycle}",
        config=cfg,
        A=A1,
        A2=A2,
        Ct=Ct,
        M=M,
        rhs=np.concatenate([f1, f2, np.zeros(m)]),  # last row of the rhs is 0 (:903)
        augment_rhs=False,
    )
    if h_scaled:
        if diagonal_inverse:
            cfg.winv_mode = b.WINV_DIAG
            prob.winv_diag = _winv_diag_from(M, squared=False)  # :705-711
        else:
            cfg.winv_mode = b.WINV_EXACT_M  # :713-720
    else:
        if diagonal_inverse:
            cfg.winv_mode = b.WINV_DIAG
            prob.winv_diag = 1.0 / (M @ M).diagonal()  # utilities.h:348-374
        else:
            cfg.winv_mode = b.WINV_EXACT_M_SQUARED
    # AMG blocks: A11 on A1 + g1 Ct diag(V) C, with V ignored (unweighted) unless the
    # diagonal inverse is in use; A22 on A2 + g2 M (:838-850)
    Vd = sp.diags(prob.winv_diag) if diagonal_inverse else sp.identity(m)
    prob.amg_matrix[b.AMG_A11] = _csr(A1 + g1 * (Ct @ Vd @ Ct.T))
    prob.amg_matrix[b.AMG_A22] = _csr(A2 + g2 * M)
    prob.amg_theta[b.AMG_A11] = 1e-3  # utilities.h:731
    prob.amg_theta[b.AMG_A22] = 1e-4
    prob.meta = dict(h=h, h_imm=h_imm, n_bg=n, m=m, cycle=cycle, beta2=beta2)
    return prob


# ----------------------------------------------------------------------------- C5: elasticity
def _elasticity_terms(lam: float, mu: float, dim: int):
    """Blocks of  2 mu eps(u):eps(v) + lam div u div v  for u = phi_j e_d, v = phi_i e_c
    (utilities.h:377-427, ElasticityUtilities::assemble_elasticity)."""

    def terms(c, d):
        if c == d:
            t = [(mu, ["K" if kk == k else "M" for kk in range(dim)]) for k in range(dim)]
            return t + [(mu + lam, ["K" if kk == c else "M" for kk in range(dim)])]
        return [
            (mu, ["D" if kk == d else ("Dt" if kk == c else "M") for kk in range(dim)]),  # d_d phi_i d_c phi_j
            (lam, ["D" if kk == c else ("Dt" if kk == d else "M") for kk in range(dim)]),  # d_c phi_i d_d phi_j
        ]

    return terms


def elasticity_interface(
    cycle: int = 1,
    lam_bg: float = 2.0,
    mu_bg: float = 1.0,
    lam_imm: float = 20.0,
    mu_imm: float = 10.0,
    gamma_fluid: float = 10.0,
    gamma_solid: float = 1e-2,
    diagonal_inverse: bool = False,
    nq_coupling: int = 3,
    nel_bg: int | None = None,
    nel_imm: int | None = None,
) -> Problem:
    """Vector-valued elliptic interface problem of parameters_elliptic_interface/elasticity.prm
    (the driver elliptic_interface_elasticity.cc is missing from the reference; the
    block structure is the 3x3 modified-AL system of elliptic_interface.cc with the blocks of
    ElasticityUtilities, utilities.h:376-589): Q1^3 on [-1.25,1.25]^3 at refinement 2+cycle,
    box inclusion (-0.65,-0.3,-0.4)-(0.65,0.3,0.4) at refinement ``cycle``.  Both spaces are
    numbered node-major (3 components per node) so the background block runs as BSR-3."""
    dim = 3
    nel = nel_bg or 2 ** (2 + cycle)
    nim = nel_imm or 2**cycle
    L = 2.5
    n1 = nel + 1
    bnd_s = boundary_mask(n1, dim)
    A1 = velocity_block_tensor(nel, dim, 0.0, bnd_s, interleaved=True, p=1, length=L,
                               terms=_elasticity_terms(lam_bg, mu_bg, dim))
    lo = np.array([-0.65, -0.3, -0.4])
    hi = np.array([0.65, 0.3, 0.4])
    # immersed tensor grid: nim cells per direction of the box (anisotropic cells)
    m1 = nim + 1
    xi = [np.linspace(lo[k], hi[k], m1) for k in range(dim)]
    xq, wq = gauss01(nq_coupling)
    pts_1d = [(xi[k][:-1, None] + (xi[k][1:] - xi[k][:-1])[:, None] * xq[None, :]).reshape(-1) for k in range(dim)]
    w_1d = [((xi[k][1:] - xi[k][:-1])[:, None] * wq[None, :]).reshape(-1) for k in range(dim)]
    Z, Y, X = np.meshgrid(pts_1d[2], pts_1d[1], pts_1d[0], indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    Wz, Wy, Wx = np.meshgrid(w_1d[2], w_1d[1], w_1d[0], indexing="ij")
    JxW = (Wx * Wy * Wz).ravel()
    Phi = background_shape_matrix(pts, nel, -1.25, 1.25, 1)
    # immersed shape functions: the same tensor-product evaluation on the box, per direction scaling
    unit = (pts - lo[None, :]) / (hi - lo)[None, :]
    Psi = background_shape_matrix(unit, nim, 0.0, 1.0, 1)
    W = sp.diags(JxW)
    Cs = zero_rows(_csr(Phi.T @ W @ Psi), bnd_s)
    Ms = _csr(Psi.T @ W @ Psi)
    I3 = sp.identity(dim, format="csr")
    Ct = _csr(sp.kron(Cs, I3, format="csr"))
    M = _csr(sp.kron(Ms, I3, format="csr"))
    # immersed stiffness with the coefficient jump, anisotropic cells: assemble per direction
    hs = (hi - lo) / nim
    K1 = [fe1d(nim, hs[k], 1, 1, 1, 1) for k in range(dim)]
    M1 = [fe1d(nim, hs[k], 1, 1) for k in range(dim)]
    D1 = [fe1d(nim, hs[k], 1, 1, 1, 0) for k in range(dim)]
    fac = {"K": K1, "M": M1, "D": D1, "Dt": [d.T.tocsr() for d in D1]}
    terms2 = _elasticity_terms(lam_imm - lam_bg, mu_imm - mu_bg, dim)
    ms = m1**dim
    rows, cols, vals = [], [], []
    for c in range(dim):
        for d in range(dim):
            blk = None
            for coef, names in terms2(c, d):
                t = coef * kron_all([fac[names[k]][k] for k in (2, 1, 0)])
                blk = t if blk is None else blk + t
            blk = blk.tocoo()
            rows.append(blk.row * dim + c)
            cols.append(blk.col * dim + d)
            vals.append(blk.data)
    A2 = _csr(sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim * ms, dim * ms)))
    n, m = Ct.shape
    h_imm = float(np.linalg.norm(hs))
    g1, g2 = gamma_fluid / h_imm**2, gamma_solid / h_imm**2
    mass_bg = kron_all([fe1d(nel, L / nel, 1, 1)] * dim) @ np.ones(n1**dim)
    f1 = np.repeat(np.where(bnd_s, 0.0, mass_bg), dim) * 1.0  # f = 1
    f2 = M @ np.ones(m) * 2.0  # f_2 = 2
    cfg = ALConfig(
        kind=b.KIND_ELLIPTIC_MODIFIED,
        restart=50,
        gamma=g1,
        gamma2=g2,
        inner=ReductionControl(10000, 1e-2, 1e-20),
        outer=ReductionControl(1000, 1e-10, 1e-6),
    )
    prob = Problem(name=f"elasticity_c{cycle}", config=cfg, A=A1, A2=A2, Ct=Ct, M=M,
                   rhs=np.concatenate([f1, f2, np.zeros(m)]), augment_rhs=False)
    if diagonal_inverse:
        cfg.winv_mode = b.WINV_DIAG
        prob.winv_diag = _winv_diag_from(M, squared=False)
    else:
        cfg.winv_mode = b.WINV_EXACT_M
    Vd = sp.diags(prob.winv_diag) if diagonal_inverse else sp.identity(m)
    prob.amg_matrix[b.AMG_A11] = _csr(A1 + g1 * (Ct @ Vd @ Ct.T))
    prob.amg_matrix[b.AMG_A22] = _csr(A2 + g2 * M)
    prob.amg_theta[b.AMG_A11] = 1e-3  # utilities.h:572-574
    prob.amg_theta[b.AMG_A22] = 1e-4
    prob.amg_comp[b.AMG_A11] = np.tile(np.arange(dim, dtype=np.int32), n1**dim)  # constant modes
    prob.amg_comp[b.AMG_A22] = np.tile(np.arange(dim, dtype=np.int32), ms)
    prob.meta = dict(h=L / nel, h_imm=h_imm, n_bg=n, m=m, cycle=cycle, node_major=True, block_size=dim)
    return prob


# ----------------------------------------------------------------------------- wiring helper
def setup_context(ctx, prob: Problem, hierarchies: dict | None = None, oracle: bool = False):
    """Hand a Problem (+ AMG hierarchies) to an ALContext / OracleContext and finalize."""
    ctx.set_csr(b.MAT_A, prob.A)
    ctx.set_csr(b.MAT_CT, prob.Ct)
    if prob.A2 is not None:
        ctx.set_csr(b.MAT_A2, prob.A2)
    if prob.Bt is not None:
        ctx.set_csr(b.MAT_BT, prob.Bt)
    if prob.Mp is not None:
        ctx.set_csr(b.MAT_MP, prob.Mp)
    ctx.set_csr(b.MAT_M, prob.M)
    if prob.winv_diag is not None:
        ctx.set_diag(b.DIAG_W_INV, prob.winv_diag)
    if oracle:
        if prob.config.winv_mode != b.WINV_DIAG:
            ctx.set_lu(0, prob.M)
        if prob.Mp is not None and prob.config.mp_inv_mode == b.MPINV_EXACT:
            ctx.set_lu(1, prob.Mp)
    if hierarchies:
        for which, H in hierarchies.items():
            ctx.set_amg(which, H)
    ctx.finalize()
    return ctx


def build_hierarchies(prob: Problem, verbose=False, **kw):
    from .amg_setup import build_hierarchy

    out = {}
    for which, Am in prob.amg_matrix.items():
        out[which] = build_hierarchy(
            Am, theta=prob.amg_theta.get(which, 1e-4), comp=prob.amg_comp.get(which), verbose=verbose, **kw
        )
    return out
