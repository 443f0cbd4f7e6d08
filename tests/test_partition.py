"""Host-side logic of the multi-GPU path (row partition, halo plans, replicated blocks),
exercised with world_size-2 `gloo` process groups on CPU: every rank builds its local
matrices with partition.py, halo values travel through torch.distributed exactly as the
plan prescribes, and the assembled results must equal the serial products."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import scipy.sparse as sp
import torch.multiprocessing as mp

from fictitious_domain_al_preconditioners_b200 import _binding as b
from fictitious_domain_al_preconditioners_b200 import partition as part
from fictitious_domain_al_preconditioners_b200 import synthetic as syn


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def halo_exchange(dc: part.DistCsr, x_owned: np.ndarray, rank: int, nranks: int) -> np.ndarray:
    """[owned | halo] vector of a DistCsr, halo filled through the plan (gloo)."""
    plan = dc.plan
    if plan is None:
        return x_owned
    soff = np.concatenate([[0], np.cumsum(plan.send_counts)])
    out = [x_owned[plan.send_idx[soff[q]: soff[q + 1]]] for q in range(nranks)]
    gathered = [None] * nranks
    dist.all_gather_object(gathered, out)
    halo = np.concatenate([gathered[q][rank] for q in range(nranks)]) if plan.n_halo else np.empty(0)
    assert halo.size == plan.n_halo
    for q in range(nranks):
        assert gathered[q][rank].size == plan.recv_counts[q]
    return np.concatenate([x_owned, halo])


def dist_spmv(dc, x_owned, rank, nranks):
    return dc.local @ halo_exchange(dc, x_owned, rank, nranks)


def allreduce(v):
    t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64).copy())
    dist.all_reduce(t)
    return t.numpy()


def dist_vcycle(LH, bvec, rank, nranks, l=0):
    """Distributed restatement of the V-cycle on LocalHierarchy pieces: levels
    below LH.rep_from are row-partitioned (halo exchanges), the restricted residual of the last
    partitioned level is all-gathered, the levels from rep_from on run replicated."""
    nl = len(LH.levels)
    rep_from = LH.rep_from if LH.rep_from >= 0 else nl
    if l == nl:
        return np.linalg.solve(LH.coarse_A.toarray(), bvec)  # replicated: bvec is the full vector
    L = LH.levels[l]
    if l >= rep_from:
        assert L.A.plan is None and L.P.plan is None and L.R.plan is None
    invd = L.inv_diag
    beta, alpha = 1.1 * L.lambda_max, L.lambda_max / LH.eig_ratio
    delta, theta = 0.5 * (beta - alpha), 0.5 * (beta + alpha)
    s1 = theta / delta
    mv = lambda v: dist_spmv(L.A, v, rank, nranks)  # noqa: E731

    def cheb(x, zero):
        rho = 1.0 / s1
        d = invd * (bvec if zero else bvec - mv(x)) / theta
        x = d.copy() if zero else x + d
        for _ in range(1, LH.cheb_degree):
            rho1 = 1.0 / (2 * s1 - rho)
            d = rho1 * rho * d + 2 * rho1 / delta * invd * (bvec - mv(x))
            x = x + d
            rho = rho1
        return x

    x = cheb(None, True)
    r = bvec - mv(x)
    rc = dist_spmv(L.R, r, rank, nranks)
    if l == rep_from - 1 and nranks > 1:
        off = LH.coarse_off
        assert rc.size == off[rank + 1] - off[rank]
        full = np.zeros(int(off[-1]))
        full[off[rank]: off[rank + 1]] = rc
        rc = allreduce(full)  # all-gather of the owned rows
        assert L.P.plan is None  # reads the replicated coarse vector through global columns
    e = dist_vcycle(LH, rc, rank, nranks, l + 1)
    x = x + dist_spmv(L.P, e, rank, nranks)
    return cheb(x, False)


def _ord(v, order):
    return v if order is None else v[order]


def _worker(rank, nranks, port, case, q):
    if case.endswith("@deep"):  # partition every level but the coarsest (no agglomeration)
        os.environ["FDAL_REP_ROWS"] = "0"
        case = case[:-5]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        _worker_body(rank, nranks, case, q)
    except Exception as e:  # surface the failure instead of a queue timeout
        import traceback

        q.put((rank, {"error": traceback.format_exc()[-1500:]}))
    finally:
        dist.destroy_process_group()


def _worker_body(rank, nranks, case, q):
    if True:
        from tests.test_oracle_known_answers import vcycle_ref

        if case == "stokes":
            prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True)
        elif case == "stokes_node":
            prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True, numbering="node")
        elif case == "stokes3d_node":
            prob = syn.stokes_immersed_boundary(dim=3, nel=6, r_emb=1, numbering="node")
        else:
            prob = syn.immersed_laplace(r_bg=4)
        H = syn.build_hierarchies(prob, max_coarse=40)
        lp = part.distribute_problem(prob, H, rank, nranks)
        rng = np.random.default_rng(3)
        X = rng.uniform(-1, 1, prob.n_dofs)
        xl = lp.scatter(X)
        n, m = prob.Ct.shape
        n0l = lp.sizes_local[0]
        x0 = xl[:n0l]
        res = {}
        # A x0 (halo), Ct lam (local), C x0 (all-reduce)
        lam = X[-m:]
        y0 = dist_spmv(lp.mats[b.MAT_A], x0, rank, nranks) + lp.mats[b.MAT_CT].local @ lam
        cx = allreduce(lp.mats[b.MAT_C].local @ x0)
        ref0 = _ord(prob.A @ X[:n] + prob.Ct @ lam, lp.order0)[lp.off0[rank]: lp.off0[rank + 1]]
        res["A"] = float(np.abs(y0 - ref0).max())
        res["C"] = float(np.abs(cx - prob.Ct.T @ X[:n]).max())
        if case.startswith("stokes"):
            n1l = lp.sizes_local[1]
            n_p = prob.Bt.shape[1]
            x1 = xl[n0l: n0l + n1l]
            yb = dist_spmv(lp.mats[b.MAT_BT], x1, rank, nranks)
            refb = _ord(prob.Bt @ X[n: n + n_p], lp.order0)[lp.off0[rank]: lp.off0[rank + 1]]
            res["Bt"] = float(np.abs(yb - refb).max())
            yB = dist_spmv(lp.mats[b.MAT_B], x0, rank, nranks)
            res["B"] = float(np.abs(yB - (prob.Bt.T @ X[:n])[lp.off1[rank]: lp.off1[rank + 1]]).max())
            ym = dist_spmv(lp.mats[b.MAT_MP], x1, rank, nranks)
            res["Mp"] = float(np.abs(ym - (prob.Mp @ X[n: n + n_p])[lp.off1[rank]: lp.off1[rank + 1]]).max())
        # Krylov dot with the replicated tail counted once
        n_dot = xl.size - (0 if rank == 0 else m)
        res["dot"] = float(abs(allreduce(np.array([xl[:n_dot] @ xl[:n_dot]]))[0] - X @ X))
        # scatter / gather round trip
        allx = [None] * nranks
        dist.all_gather_object(allx, xl)
        res["roundtrip"] = float(np.abs(lp.gather(allx) - X).max())
        # distributed V-cycle == serial V-cycle
        LH = lp.amg[b.AMG_A11]
        z = dist_vcycle(LH, x0, rank, nranks)
        zref = _ord(vcycle_ref(H[b.AMG_A11], X[:n]), lp.order0)[lp.off0[rank]: lp.off0[rank + 1]]
        res["vcycle"] = float(np.abs(z - zref).max() / np.abs(zref).max())
        res["levels"] = len(LH.levels)
        res["rep_from"] = LH.rep_from
        q.put((rank, res))


@pytest.mark.parametrize("case,nranks", [("laplace", 2), ("stokes", 2), ("stokes_node", 2), ("stokes3d_node", 4),
                                         ("laplace@deep", 2), ("stokes_node@deep", 2), ("stokes3d_node@deep", 4)])
def test_partition_matches_serial(case, nranks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, nranks, port, case, q)) for r in range(nranks)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert "error" not in res, res.get("error")
        nlev = res.pop("levels")
        rep = res.pop("rep_from")
        assert nlev >= 1 and 1 <= rep <= nlev
        if case.endswith("@deep"):
            assert rep == nlev
        for k, v in res.items():
            assert v < 1e-11, (rank, k, v)


def test_split_offsets_keeps_nodes_together():
    off = part.split_offsets(2 * 17 * 17, 3, align=2)
    assert off[0] == 0 and off[-1] == 2 * 17 * 17
    assert all(o % 2 == 0 for o in off)
    assert np.all(np.diff(off) > 0)


def test_halo_plan_is_consistent_single_process():
    """send lists of rank q towards r hold exactly the halo columns r expects from q."""
    prob = syn.immersed_laplace(r_bg=4)
    n = prob.A.shape[0]
    P = 4
    off = part.split_offsets(n, P)
    dcs = [part.localize(prob.A, off, off, r) for r in range(P)]
    for r in range(P):
        roff = np.concatenate([[0], np.cumsum(dcs[r].plan.recv_counts)])
        for q in range(P):
            soff = np.concatenate([[0], np.cumsum(dcs[q].plan.send_counts)])
            sent = dcs[q].plan.send_idx[soff[r]: soff[r + 1]] + off[q]
            want = dcs[r].plan.halo_globals[roff[q]: roff[q + 1]]
            assert np.array_equal(sent, want)


def test_coarse_ownership_depends_on_the_pattern_only():
    """Every rank cuts its own copy of the hierarchy: ownership must not move when the
    floating-point values differ in the last bits (an arg-max rule has cross-rank ties and
    dead-locked the 4-GPU halo exchange in round 1)."""
    prob = syn.stokes_immersed_boundary(dim=3, nel=8, numbering="node")
    H = syn.build_hierarchies(prob, max_coarse=300)
    P0 = H[b.AMG_A11].levels[0].P
    off = part.split_offsets(P0.shape[0], 4, 3)
    rng = np.random.default_rng(0)
    P1 = P0.copy()
    P1.data = P1.data * (1.0 + 1e-13 * rng.standard_normal(P1.data.size))
    o0, f0 = part.coarse_order(P0, off)
    o1, f1 = part.coarse_order(P1, off)
    assert np.array_equal(o0, o1) and np.array_equal(f0, f1)
    assert f0[-1] == P0.shape[1] and np.all(np.diff(f0) >= 0)


def _share_worker(rank, nranks, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        built = []

        def build():
            built.append(rank)
            prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True, numbering="node")
            return prob, syn.build_hierarchies(prob, max_coarse=40), {"n_dofs": prob.n_dofs}

        lp = part.share_local_problems(build, rank, nranks)
        prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True, numbering="node")
        ref = part.distribute_problem(prob, syn.build_hierarchies(prob, max_coarse=40), rank, nranks)
        ok = built == ([0] if rank == 0 else [])  # only rank 0 ran the setup
        ok &= lp.rank == rank and lp.meta["n_dofs"] == prob.n_dofs and lp.augment_rhs
        ok &= np.array_equal(lp.rhs_local, ref.scatter(prob.rhs))
        for mid in ref.mats:
            ok &= (lp.mats[mid].local != ref.mats[mid].local).nnz == 0
            if ref.mats[mid].plan is not None:
                ok &= np.array_equal(lp.mats[mid].plan.send_idx, ref.mats[mid].plan.send_idx)
        ok &= len(lp.amg[b.AMG_A11].levels) == len(ref.amg[b.AMG_A11].levels)
        q.put((rank, bool(ok)))
    except Exception:
        import traceback

        q.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def test_rank0_setup_is_shared_with_the_other_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_share_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok in out:
        assert ok is True, ok


def test_single_rank_share_is_the_renumbered_problem():
    """nranks = 1 (the BSR path of bench.py and the GPU tests): the local problem is the global
    one, renumbered node-major, with empty halo plans and an unchanged hierarchy."""
    for numbering in ("node", "component"):
        prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True, numbering=numbering)
        H = syn.build_hierarchies(prob, max_coarse=40)
        lp = part.distribute_problem(prob, H, 0, 1)
        n = prob.A.shape[0]
        order = lp.order0 if lp.order0 is not None else np.arange(n)
        assert lp.block_size == 2 and lp.sizes_local == prob.sizes
        assert (lp.mats[b.MAT_A].local != prob.A[order][:, order]).nnz == 0
        assert (lp.mats[b.MAT_CT].local != prob.Ct[order]).nnz == 0
        assert (lp.mats[b.MAT_C].local != sp.csr_matrix(prob.Ct[order].T)).nnz == 0
        assert (lp.mats[b.MAT_BT].local != prob.Bt[order]).nnz == 0
        assert (lp.mats[b.MAT_B].local != sp.csr_matrix(prob.Bt[order].T)).nnz == 0
        assert (lp.mats[b.MAT_MP].local != prob.Mp).nnz == 0
        for dc in lp.mats.values():
            assert dc.plan is None or (dc.plan.n_halo == 0 and dc.plan.send_idx.size == 0)
        LH = lp.amg[b.AMG_A11]
        assert len(LH.levels) == len(H[b.AMG_A11].levels) - 1
        L0, G0 = LH.levels[0], H[b.AMG_A11].levels[0]
        assert (L0.A.local != G0.A[order][:, order]).nnz == 0
        assert (L0.P.local != G0.P[order]).nnz == 0 and (L0.R.local != G0.R[:, order]).nnz == 0
        assert np.array_equal(L0.inv_diag, G0.inv_diag[order]) and L0.lambda_max == G0.lambda_max
        assert (LH.coarse_A != H[b.AMG_A11].levels[-1].A).nnz == 0
        assert list(LH.coarse_off) == [0, LH.coarse_A.shape[0]]
        x = np.random.default_rng(0).uniform(-1, 1, prob.n_dofs)
        assert np.array_equal(lp.gather([lp.scatter(x)]), x)


def test_inconsistent_halo_plans_are_refused_on_the_host():
    """A plan in which one rank sends what its peer does not expect would deadlock the grouped
    send/recv on the GPUs; check_plan_signatures turns it into an error before any context exists."""
    prob = syn.stokes_immersed_boundary(dim=2, nel=8, diagonal_mass=True, numbering="node")
    H = syn.build_hierarchies(prob, max_coarse=40)
    lps = part.distribute_all(prob, H, 3)  # validates internally
    sigs = [part.plan_signature(lp) for lp in lps]
    assert any(k[0] == "amg" for k in sigs[0]) and ("mat", b.MAT_A) in sigs[0]
    part.check_plan_signatures(sigs)
    sigs[1][("mat", b.MAT_A)][0][0] += 1
    with pytest.raises(ValueError, match="halo plan mismatch"):
        part.check_plan_signatures(sigs)
    # problems with a replicated immersed-block hierarchy (elliptic kinds) pass through the check too
    prob = syn.elliptic_interface(cycle=1)
    lps = part.distribute_all(prob, syn.build_hierarchies(prob, max_coarse=40), 2)
    assert not isinstance(lps[0].amg[b.AMG_A22], part.LocalHierarchy) and part.plan_signature(lps[0])


def _emulated_spmv(A, row_off, col_off, x, col_bs):
    """What the ranks compute together: pack -> exchange (by the plans) -> local SpMV on [owned | halo]."""
    nranks = len(row_off) - 1
    part.clear_cache()
    dcs = [part.localize(A, row_off, col_off, r, col_bs) for r in range(nranks)]
    part.clear_cache()
    owned = [x[col_off[r]: col_off[r + 1]] for r in range(nranks)]
    out = []
    for r in range(nranks):
        pl = dcs[r].plan
        halo = []
        for q in range(nranks):  # halo entries are ordered by owner rank
            so = np.concatenate([[0], np.cumsum(dcs[q].plan.send_counts)])
            sent = owned[q][dcs[q].plan.send_idx[so[r]: so[r + 1]]]
            assert sent.size == pl.recv_counts[q]
            halo.append(sent)
        xl = np.concatenate([owned[r]] + halo)
        assert xl.size == dcs[r].local.shape[1] == pl.n_owned + pl.n_halo
        out.append(dcs[r].local @ xl)
    return np.concatenate(out)


@pytest.mark.parametrize("seed", range(6))
def test_random_matrices_through_emulated_halo_exchange(seed):
    """Random sparse (also rectangular) matrices, random rank counts including ranks that own no rows
    or no columns, scalar and node-interleaved column spaces: the plans reproduce the global product."""
    rng = np.random.default_rng(seed)
    bs = int(rng.choice([1, 1, 2, 3]))
    nranks = int(rng.integers(2, 7))
    nr = int(rng.integers(1, 40)) * bs
    nc = int(rng.integers(1, 40)) * bs
    A = sp.random(nr, nc, density=float(rng.uniform(0.02, 0.4)), random_state=seed, format="csr")
    x = rng.uniform(-1, 1, nc)

    def offsets(n):  # random cuts on node boundaries, empty shares allowed
        cuts = np.sort(rng.integers(0, n // bs + 1, nranks - 1)) * bs
        return np.concatenate([[0], cuts, [n]]).astype(np.int64)

    row_off, col_off = offsets(nr), offsets(nc)
    y = _emulated_spmv(A, row_off, col_off, x, bs)
    assert np.allclose(y, A @ x, rtol=0, atol=1e-13)
    if bs > 1:  # halos hold whole nodes
        part.clear_cache()
        for r in range(nranks):
            hg = part.localize(A, row_off, col_off, r, bs).plan.halo_globals
            assert hg.size % bs == 0 and np.array_equal(hg.reshape(-1, bs)[:, 0] % bs, np.zeros(hg.size // bs))
        part.clear_cache()


def test_openmp_cut_helpers_match_the_numpy_path(monkeypatch):
    """localize() cuts the 10^9-entry fine matrices with the OpenMP helpers of csrc/host_setup.c; on the same
    matrix they must give exactly the numpy path's local matrices and halo plans."""
    prob = syn.stokes_immersed_boundary(dim=3, nel=6, r_emb=1, numbering="node")
    A = prob.A.tocsr()
    off = part.split_offsets(A.shape[0], 4, 3)
    ref = [part.localize(A, off, off, r, 3) for r in range(4)]
    part.clear_cache()
    monkeypatch.setattr(part, "_C_PATH_MIN_NNZ", 0)
    got = [part.localize(A, off, off, r, 3) for r in range(4)]
    part.clear_cache()
    for a, g in zip(ref, got):
        assert (a.local != g.local).nnz == 0 and a.local.shape == g.local.shape
        assert np.array_equal(a.local.indices, g.local.indices) and np.array_equal(a.local.indptr, g.local.indptr)
        for f in ("send_counts", "send_idx", "recv_counts", "halo_globals"):
            assert np.array_equal(getattr(a.plan, f), getattr(g.plan, f)), f
