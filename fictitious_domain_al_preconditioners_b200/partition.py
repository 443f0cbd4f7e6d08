"""Row partition of the AL solve path over the GPUs of one box (host, setup only).

The reference is serial; this decomposition is new.  Rows of the
background unknowns (and, for Stokes, of the pressure unknowns) are split into
contiguous ranges after a locality-preserving renumbering; every matrix with those
rows is split the same way; the multiplier block (m <= ~10^5) and the immersed blocks
are replicated on every rank.  A row-partitioned matrix numbers its columns
``[owned | halo]`` and carries a halo plan (who sends which owned entries to whom);
the CUDA library packs, exchanges (NCCL send/recv) and gathers through that plan.

  SpMV(A, Bt, B, Mp, AMG A_l/P_l/R_l) : halo exchange of the column space, local rows
  C x                                  : local partial over owned columns + all-reduce(m)
  Ct t                                 : purely local (t replicated)
  Krylov dots                          : local partial + all-reduce; replicated tail
                                         blocks are counted on rank 0 only
  coarsest AMG level                   : right-hand side all-reduced into a replicated
                                         vector, every rank applies its rows of the inverse
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import _binding as b


def split_offsets(n: int, nranks: int, align: int = 1) -> np.ndarray:
    """Contiguous, nearly equal ranges; boundaries are multiples of `align`
    (keeps the components of one node on one rank)."""
    units = n // align
    base, rem = divmod(units, nranks)
    counts = np.array([base + (1 if r < rem else 0) for r in range(nranks)], dtype=np.int64) * align
    counts[-1] += n - counts.sum()
    return np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)


@dataclass
class HaloPlan:
    n_owned: int
    n_halo: int
    send_counts: np.ndarray  # int32[nranks]
    send_idx: np.ndarray  # int32[sum(send_counts)], local owned indices, grouped by destination rank
    recv_counts: np.ndarray  # int32[nranks]; halo entries are ordered by owner rank, then global index
    halo_globals: np.ndarray  # global (renumbered) column of each halo entry (for tests)


@dataclass
class DistCsr:
    local: sp.csr_matrix  # rows: owned rows; columns: [owned | halo]
    plan: HaloPlan | None  # None: column space is replicated / fully local


def _node_complete(cols: np.ndarray, bs: int) -> np.ndarray:
    """All components of every node touched by `cols` (BSR needs whole nodes in the halo)."""
    if bs <= 1:
        return np.unique(cols)
    nodes = np.unique(cols // bs)
    return (nodes[:, None] * bs + np.arange(bs, dtype=nodes.dtype)[None, :]).reshape(-1)


_NEEDS_CACHE: dict = {}
_C_PATH_MIN_NNZ = 1_000_000  # row slices with more entries than this are cut by the OpenMP helpers


def clear_cache():
    _NEEDS_CACHE.clear()


def _host_lib():
    import ctypes as C

    from .amg_setup import _host

    lib = _host()
    if not getattr(lib, "_part_ready", False):
        p64, p32, pu8 = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        lib.fdal_host_mark_foreign_cols.restype = None
        lib.fdal_host_mark_foreign_cols.argtypes = [C.c_int64, p32, C.c_int64, C.c_int64, pu8]
        lib.fdal_host_localize_cols.restype = None
        lib.fdal_host_localize_cols.argtypes = [C.c_int64, p32, C.c_int64, C.c_int64, p64, C.c_int64, p32]
        lib._part_ready = True
    return lib


def _halo_needs(A: sp.csr_matrix, row_off, col_off, col_bs: int):
    """For every rank q: the sorted, node-complete list of columns its rows touch outside its
    own column range.  Cached per matrix object so that cutting all ranks' shares on one
    process (rank-0 setup) costs one pass instead of one per rank.  The scan over the 10^9 column
    indices of the fine matrices runs in the OpenMP helper (csrc/host_setup.c)."""
    import ctypes as C

    key = (id(A), A.nnz, tuple(int(x) for x in row_off), tuple(int(x) for x in col_off), col_bs)
    hit = _NEEDS_CACHE.get(key)
    if hit is not None:
        return hit
    nranks = len(row_off) - 1
    lib = _host_lib() if A.indices.dtype == np.int32 else None
    out = []
    for q in range(nranks):
        lo, hi = int(A.indptr[int(row_off[q])]), int(A.indptr[int(row_off[q + 1])])
        q0, q1 = int(col_off[q]), int(col_off[q + 1])
        if lib is not None and hi - lo > _C_PATH_MIN_NNZ:
            marks = np.zeros(A.shape[1], dtype=np.uint8)
            idx = A.indices[lo:hi]
            lib.fdal_host_mark_foreign_cols(hi - lo, idx.ctypes.data_as(C.POINTER(C.c_int32)), q0, q1,
                                            marks.ctypes.data_as(C.POINTER(C.c_uint8)))
            foreign = np.flatnonzero(marks).astype(np.int64)
            out.append(_node_complete(foreign, col_bs))
        else:
            cq = A.indices[lo:hi].astype(np.int64)
            out.append(_node_complete(cq[(cq < q0) | (cq >= q1)], col_bs))
    _NEEDS_CACHE[key] = out
    return out


def localize(A: sp.csr_matrix, row_off: np.ndarray, col_off: np.ndarray, rank: int, col_bs: int = 1) -> DistCsr:
    """Rows row_off[rank]:row_off[rank+1] of the (renumbered) global matrix with a halo plan
    over the column space partitioned by col_off.  col_bs > 1: the column space is
    node-interleaved with col_bs components per node and halos hold whole nodes."""
    import ctypes as C

    nranks = len(row_off) - 1
    A = A.tocsr()
    if nranks == 1:  # nothing to cut: the whole matrix, empty plan
        z = np.zeros(1, dtype=np.int32)
        return DistCsr(A, HaloPlan(A.shape[1], 0, z, np.empty(0, np.int32), z.copy(), np.empty(0, np.int64)))
    c0, c1 = int(col_off[rank]), int(col_off[rank + 1])
    r0, r1 = int(row_off[rank]), int(row_off[rank + 1])
    needs = _halo_needs(A, row_off, col_off, col_bs)  # per rank: node-complete non-owned columns
    halo_globals = np.ascontiguousarray(needs[rank], dtype=np.int64)
    lo, hi = int(A.indptr[r0]), int(A.indptr[r1])
    indptr = (A.indptr[r0: r1 + 1] - A.indptr[r0]).astype(np.int64 if hi - lo >= 2**31 - 1 else np.int32)
    if A.indices.dtype == np.int32 and hi - lo > _C_PATH_MIN_NNZ:
        # views of the owner's rows (no copy of the values), columns renumbered by the OpenMP helper
        idx = A.indices[lo:hi]
        newcol = np.empty(hi - lo, dtype=np.int32)
        _host_lib().fdal_host_localize_cols(hi - lo, idx.ctypes.data_as(C.POINTER(C.c_int32)), c0, c1,
                                            halo_globals.ctypes.data_as(C.POINTER(C.c_int64)), halo_globals.size,
                                            newcol.ctypes.data_as(C.POINTER(C.c_int32)))
        data = A.data[lo:hi]
    else:
        cols = A.indices[lo:hi].astype(np.int64)
        owned = (cols >= c0) & (cols < c1)
        newcol = np.empty(cols.size, dtype=np.int32)
        newcol[owned] = cols[owned] - c0
        newcol[~owned] = (c1 - c0) + np.searchsorted(halo_globals, cols[~owned])
        data = A.data[lo:hi]
    local = sp.csr_matrix((data, newcol, indptr), shape=(r1 - r0, (c1 - c0) + halo_globals.size), copy=False)
    owner = np.searchsorted(col_off, halo_globals, side="right") - 1
    recv_counts = np.bincount(owner, minlength=nranks).astype(np.int32)
    # what every other rank needs from my owned range
    send_lists = []
    for q in range(nranks):
        if q == rank:
            send_lists.append(np.empty(0, dtype=np.int64))
            continue
        cq = needs[q]
        mine = cq[(cq >= c0) & (cq < c1)]
        send_lists.append(mine - c0)
    send_counts = np.array([s.size for s in send_lists], dtype=np.int32)
    send_idx = np.concatenate(send_lists).astype(np.int32) if send_lists else np.empty(0, np.int32)
    plan = HaloPlan(c1 - c0, int(halo_globals.size), send_counts, send_idx, recv_counts, halo_globals)
    return DistCsr(local, plan)


def block0_order(prob) -> tuple[np.ndarray, int]:
    """new->old order of the background block and the node block size.  Vector-valued
    blocks numbered component-wise ([u_x|u_y|u_z], DoFRenumbering::component_wise) are
    interleaved node-major so that a contiguous range is a spatial slab holding all
    components of its nodes."""
    n = prob.A.shape[0]
    if prob.meta.get("node_major"):
        return None, int(prob.meta["block_size"])  # already node-interleaved: identity
    comp = prob.amg_comp.get(b.AMG_A11)
    if comp is None:
        return None, 1
    dim = int(comp.max()) + 1
    ns = n // dim
    node = np.arange(ns, dtype=np.int64)
    order = (node[:, None] + ns * np.arange(dim, dtype=np.int64)[None, :]).reshape(-1)
    return order, dim


def coarse_order(P: sp.csr_matrix, fine_off: np.ndarray):
    """Ownership of the coarse unknowns, from the sparsity PATTERN of P only: a coarse unknown
    lives with the median fine row of its column.  Every rank derives the partition from its
    own copy of the hierarchy, so the rule must not depend on floating-point values (an
    arg-max over |P_ij| has ties that round differently when the setup ran on different
    devices, which would give the ranks inconsistent halo plans and dead-lock the exchange).
    Returns new->old order and the offsets."""
    nranks = len(fine_off) - 1
    if nranks == 1:
        return None, np.array([0, P.shape[1]], dtype=np.int64)
    Pc = P.tocsc()
    Pc.sort_indices()
    nc = Pc.shape[1]
    cnt = np.diff(Pc.indptr)
    has = cnt > 0
    mid = Pc.indptr[:-1] + cnt // 2
    rows_mid = np.zeros(nc, dtype=np.int64)
    rows_mid[has] = Pc.indices[mid[has]]
    owner = np.searchsorted(fine_off, rows_mid, side="right") - 1
    owner[~has] = 0
    order = np.lexsort((np.arange(nc), owner)).astype(np.int64)
    counts = np.bincount(owner, minlength=nranks)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return order, off


@dataclass
class LocalLevel:
    A: DistCsr
    P: DistCsr | None
    R: DistCsr | None
    inv_diag: np.ndarray | None
    lambda_max: float


@dataclass
class LocalHierarchy:
    levels: list  # levels[:rep_from] row-partitioned, levels[rep_from:] replicated (all but the coarsest)
    coarse_A: sp.csr_matrix  # replicated coarsest operator
    coarse_off: np.ndarray  # ownership of the rows of level rep_from (all-gather of the restricted residual)
    cheb_degree: int
    eig_ratio: float
    rep_from: int = -1  # first replicated level (-1: only the coarsest operator)
    replicated: bool = False


def replicate_from(nnzs, nranks: int, max_nnz: int | None = None) -> int:
    """First level of the hierarchy that is agglomerated onto every rank: the first one (after the
    finest) whose operator has at most ``max_nnz`` non-zeros; the coarsest level is always replicated.
    Rule: a partitioned mat-vec costs t/N + ~45 us (push kernel, launch, waiting for the slowest
    neighbour; measured on 2 and 4 B200s), a replicated one t = 12 B * nnz / ~5 TB/s, so replication
    wins while nnz <= 1.9e7 / (1 - 1/N) (env FDAL_REP_NNZ overrides; FDAL_REP_ROWS=0 keeps every level
    but the coarsest partitioned).  The rule looks at non-zeros, not rows: the coarse operators of the
    3-D Q2 problems have few rows but 600-1400 entries per row."""
    import os

    if max_nnz is None:
        if os.environ.get("FDAL_REP_ROWS") == "0":
            max_nnz = 0
        elif os.environ.get("FDAL_REP_NNZ"):
            max_nnz = int(float(os.environ["FDAL_REP_NNZ"]))
        else:
            max_nnz = int(1.9e7 / (1.0 - 1.0 / max(2, nranks)))
    nl = len(nnzs)
    for l in range(1, nl):
        if nnzs[l] <= max_nnz:
            return l
    return nl - 1


@dataclass
class LocalProblem:
    rank: int
    nranks: int
    config: object
    mats: dict = field(default_factory=dict)  # matrix id -> DistCsr
    amg: dict = field(default_factory=dict)  # which -> LocalHierarchy
    winv_diag: np.ndarray | None = None
    off0: np.ndarray | None = None
    off1: np.ndarray | None = None
    order0: np.ndarray | None = None
    order1: np.ndarray | None = None
    sizes_global: tuple = ()
    sizes_local: tuple = ()
    block_size: int = 1
    rhs_local: np.ndarray | None = None  # filled by share_local_problems
    augment_rhs: bool = False
    meta: dict = field(default_factory=dict)

    # ---- vectors ---------------------------------------------------------------------
    def scatter(self, x_global: np.ndarray) -> np.ndarray:
        """Global block vector -> this rank's local vector [block0_loc | block1_loc/rep | rep]."""
        n0 = self.sizes_global[0]
        x0 = x_global[:n0] if self.order0 is None else x_global[:n0][self.order0]
        parts = [x0[self.off0[self.rank]: self.off0[self.rank + 1]]]
        rest = x_global[n0:]
        if self.off1 is not None:
            n1 = self.sizes_global[1]
            parts.append(rest[:n1][self.order1][self.off1[self.rank]: self.off1[self.rank + 1]])
            parts.append(rest[n1:])
        else:
            parts.append(rest)
        return np.concatenate(parts)

    def gather(self, locals_: list) -> np.ndarray:
        """Inverse of scatter given the local vectors of all ranks (tests / output)."""
        n0 = self.sizes_global[0]
        out = np.zeros(sum(self.sizes_global))
        b0 = np.concatenate([v[: self.off0[r + 1] - self.off0[r]] for r, v in enumerate(locals_)])
        if self.order0 is None:
            out[:n0] = b0
        else:
            out[:n0][self.order0] = b0
        if self.off1 is not None:
            n1 = self.sizes_global[1]
            b1 = np.concatenate([
                v[self.off0[r + 1] - self.off0[r]: self.off0[r + 1] - self.off0[r] + self.off1[r + 1] - self.off1[r]]
                for r, v in enumerate(locals_)])
            tmp = np.zeros(n1)
            tmp[self.order1] = b1
            out[n0:n0 + n1] = tmp
            out[n0 + n1:] = locals_[0][(self.off0[1] - self.off0[0]) + (self.off1[1] - self.off1[0]):]
        else:
            out[n0:] = locals_[0][self.off0[1] - self.off0[0]:]
        return out


def _perm(A, row_order, col_order):
    A = A.tocsr()
    if row_order is None and col_order is None:
        return A
    if row_order is not None:
        A = A[row_order]
    if col_order is not None:
        A = A[:, col_order]
    A = A.tocsr()
    A.sort_indices()
    return A


class Distributor:
    """Cuts a Problem (+ AMG hierarchies) for `nranks` ranks.  Everything that does not depend
    on the rank — renumbered matrices, transposes, coarse ownership — is computed once, so
    cutting all ranks' shares on one process (rank-0 setup) is one pass plus row slicing."""

    def __init__(self, prob, hierarchies: dict, nranks: int):
        self.prob, self.nranks = prob, nranks
        kind = prob.config.kind
        n, m = prob.Ct.shape
        self.order0, self.bs = block0_order(prob)
        self.off0 = split_offsets(n, nranks, self.bs)
        self.A = _perm(prob.A, self.order0, self.order0)
        self.Ct = _perm(prob.Ct, self.order0, None)
        self.C = sp.csr_matrix(self.Ct.T).tocsc()  # column slices per rank
        self.stokes = kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES)
        if self.stokes:
            n_p = prob.Bt.shape[1]
            self.order1 = np.arange(n_p, dtype=np.int64)
            self.off1 = split_offsets(n_p, nranks)
            self.Bt = _perm(prob.Bt, self.order0, self.order1)
            self.B = sp.csr_matrix(self.Bt.T)
            self.Mp = _perm(prob.Mp, self.order1, self.order1)
        # hierarchy of the background block: renumber every level once
        self.hier = {}
        for which, H in hierarchies.items():
            if which != b.AMG_A11:
                self.hier[which] = H  # immersed block: replicated
                continue
            nl = len(H.levels)
            rep_from = replicate_from([L.A.nnz for L in H.levels], nranks) if nranks > 1 else nl - 1
            levels = []
            order_f, off_f = self.order0, self.off0
            rep_off = off_f
            Aperm = _perm(H.levels[0].A, order_f, order_f)
            for l in range(nl - 1):
                L = H.levels[l]
                if l < rep_from:
                    order_c, off_c = coarse_order(_perm(L.P, order_f, None), off_f)
                else:
                    order_c, off_c = None, None  # replicated levels keep the hierarchy's own numbering
                Pp = _perm(L.P, order_f, order_c)
                Rp = _perm(L.R if L.R is not None else L.P.T, order_c, order_f)
                invd = None if L.inv_diag is None else (L.inv_diag if order_f is None else L.inv_diag[order_f])
                levels.append(dict(A=Aperm, P=Pp, R=Rp, inv_diag=invd, lambda_max=L.lambda_max, off_f=off_f, off_c=off_c,
                                   bs=self.bs if l == 0 else 1, part=l < rep_from, last_part=l == rep_from - 1))
                if l == rep_from - 1:
                    rep_off = off_c
                order_f, off_f = order_c, off_c
                Aperm = _perm(H.levels[l + 1].A, order_f, order_f)
            self.hier[which] = dict(levels=levels, coarse_A=Aperm, coarse_off=rep_off, cheb_degree=H.cheb_degree,
                                    eig_ratio=H.eig_ratio, rep_from=rep_from)

    def local(self, rank: int) -> LocalProblem:
        prob, nranks, off0, bs = self.prob, self.nranks, self.off0, self.bs
        kind = prob.config.kind
        m = prob.Ct.shape[1]
        lp = LocalProblem(rank=rank, nranks=nranks, config=prob.config, order0=self.order0, off0=off0,
                          sizes_global=prob.sizes, winv_diag=prob.winv_diag)
        lp.block_size = bs
        r0, r1 = int(off0[rank]), int(off0[rank + 1])
        lp.mats[b.MAT_A] = localize(self.A, off0, off0, rank, bs)
        lp.mats[b.MAT_CT] = DistCsr(self.Ct[r0:r1].tocsr(), None)
        lp.mats[b.MAT_C] = DistCsr(self.C[:, r0:r1].tocsr(), None)
        lp.mats[b.MAT_M] = DistCsr(prob.M.tocsr(), None)
        n_loc = r1 - r0
        if self.stokes:
            off1 = self.off1
            lp.order1, lp.off1 = self.order1, off1
            lp.mats[b.MAT_BT] = localize(self.Bt, off0, off1, rank)
            lp.mats[b.MAT_B] = localize(self.B, off1, off0, rank, bs)
            lp.mats[b.MAT_MP] = localize(self.Mp, off1, off1, rank)
            lp.sizes_local = (n_loc, int(off1[rank + 1] - off1[rank]), m)
        elif kind == b.KIND_LAPLACE:
            lp.sizes_local = (n_loc, m)
        else:
            lp.mats[b.MAT_A2] = DistCsr(prob.A2.tocsr(), None)
            lp.sizes_local = (n_loc, m, m)
        for which, Hd in self.hier.items():
            if which != b.AMG_A11:
                lp.amg[which] = Hd
                continue
            levels = []
            for L in Hd["levels"]:
                off_f, off_c = L["off_f"], L["off_c"]
                if not L["part"]:  # agglomerated level: the whole matrices on every rank, no exchange
                    levels.append(LocalLevel(A=DistCsr(L["A"], None), P=DistCsr(L["P"], None), R=DistCsr(L["R"], None),
                                             inv_diag=L["inv_diag"], lambda_max=L["lambda_max"]))
                    continue
                if L["last_part"] and nranks > 1:
                    # the next level is replicated: P reads the full coarse vector (global columns, no
                    # halo), R produces this rank's rows of it (all-gathered afterwards)
                    P_loc = DistCsr(L["P"][int(off_f[rank]): int(off_f[rank + 1])].tocsr(), None)
                else:
                    P_loc = localize(L["P"], off_f, off_c, rank)
                levels.append(LocalLevel(
                    A=localize(L["A"], off_f, off_f, rank, L["bs"]),
                    P=P_loc,
                    R=localize(L["R"], off_c, off_f, rank, L["bs"]),
                    inv_diag=None if L["inv_diag"] is None else L["inv_diag"][off_f[rank]: off_f[rank + 1]],
                    lambda_max=L["lambda_max"]))
            lp.amg[which] = LocalHierarchy(levels=levels, coarse_A=Hd["coarse_A"], coarse_off=Hd["coarse_off"],
                                           cheb_degree=Hd["cheb_degree"], eig_ratio=Hd["eig_ratio"],
                                           rep_from=Hd["rep_from"])
        return lp


def distribute_problem(prob, hierarchies: dict, rank: int, nranks: int) -> LocalProblem:
    """This rank's share of a Problem (+ AMG hierarchies)."""
    lp = Distributor(prob, hierarchies, nranks).local(rank)
    clear_cache()
    return lp


def setup_local_context(ctx, lp: LocalProblem, uid: bytes = bytes(128)):
    """Hand this rank's LocalProblem to its ALContext (after ``comm_init``) and finalize.
    With a single rank this is simply the node-interleaved (BSR-ready) renumbering of the
    problem; ``ctx.config.block_size`` should then be ``lp.block_size``."""
    ctx.comm_init(uid, lp.rank, lp.nranks)
    for mid, dc in lp.mats.items():
        ctx.set_csr(mid, dc.local)
        if dc.plan is not None:
            ctx.set_halo(mid, dc.plan)
    if lp.winv_diag is not None:
        ctx.set_diag(b.DIAG_W_INV, lp.winv_diag)
    for which, H in lp.amg.items():
        if isinstance(H, LocalHierarchy):
            ctx.set_amg_local(which, H)
        else:
            ctx.set_amg(which, H)
    ctx.finalize()
    return ctx


def plan_signature(lp: LocalProblem) -> dict:
    """(send_counts, recv_counts) of every halo plan of one rank, keyed by matrix."""
    sig = {}
    for mid, dc in lp.mats.items():
        if dc.plan is not None:
            sig[("mat", mid)] = (dc.plan.send_counts.copy(), dc.plan.recv_counts.copy())
    for which, LH in lp.amg.items():
        if not isinstance(LH, LocalHierarchy):
            continue  # replicated hierarchy of an immersed block: no exchange
        for l, L in enumerate(LH.levels):
            for tag, dc in (("A", L.A), ("P", L.P), ("R", L.R)):
                if dc is not None and dc.plan is not None:
                    sig[("amg", which, l, tag)] = (dc.plan.send_counts.copy(), dc.plan.recv_counts.copy())
    return sig


def check_plan_signatures(sigs: list) -> None:
    """What rank p sends to q is what q expects from p, for every partitioned matrix.  A mismatch
    would leave the grouped ncclSend/ncclRecv of halo_exchange waiting forever, so it is refused on
    the host before any context is built."""
    nranks = len(sigs)
    keys = set(sigs[0])
    for r, sg in enumerate(sigs):
        if set(sg) != keys:
            raise ValueError(f"rank {r} has halo plans for {sorted(set(sg) ^ keys)} that rank 0 lacks (or vice versa)")
    for key in sorted(keys, key=str):
        for p_ in range(nranks):
            send = sigs[p_][key][0]
            for q in range(nranks):
                if int(send[q]) != int(sigs[q][key][1][p_]):
                    raise ValueError(f"halo plan mismatch for {key}: rank {p_} sends {int(send[q])} entries to rank {q}, "
                                     f"which expects {int(sigs[q][key][1][p_])}")
            if int(send[p_]) != 0:
                raise ValueError(f"halo plan for {key}: rank {p_} sends to itself")


def distribute_all(prob, hierarchies: dict, nranks: int) -> list:
    """Every rank's LocalProblem, cut on one process."""
    D = Distributor(prob, hierarchies, nranks)
    out = [D.local(r) for r in range(nranks)]
    clear_cache()
    check_plan_signatures([plan_signature(lp) for lp in out])
    return out


def share_local_problems(build_fn, rank: int, nranks: int, group=None):
    """Setup on rank 0 only: ``build_fn()`` -> (problem, hierarchies, meta dict) runs once, the
    ranks' shares travel through pickle files in shared memory.  One copy of the global
    problem in host memory instead of one per rank, and every rank's plan comes from the SAME
    hierarchy (nothing to disagree about).  Returns this rank's LocalProblem with ``rhs_local``,
    ``augment_rhs`` and ``meta`` filled in."""
    import os
    import pickle
    import shutil
    import tempfile

    import torch.distributed as dist

    box = [None]
    if rank == 0:
        try:
            prob, H, meta = build_fn()
            # shared memory when it is large enough (container /dev/shm is often 64 MB), else /tmp
            cands = [p_ for p_ in ("/dev/shm", tempfile.gettempdir()) if os.path.isdir(p_) and os.access(p_, os.W_OK)]
            base = max(cands, key=lambda p_: shutil.disk_usage(p_).free) if cands else None
            d = tempfile.mkdtemp(prefix="fdal_lp_", dir=base)
            D = Distributor(prob, H, nranks)
            sigs = []
            for r in range(nranks):
                lp = D.local(r)
                sigs.append(plan_signature(lp))
                lp.rhs_local = lp.scatter(prob.rhs)
                lp.augment_rhs = bool(prob.augment_rhs)
                lp.meta = dict(meta)
                with open(os.path.join(d, f"lp_{r}.pkl"), "wb") as f:
                    pickle.dump(lp, f, protocol=5)
                del lp
            clear_cache()
            del prob, H, D
            check_plan_signatures(sigs)
            box[0] = d
        except Exception as e:  # tell the other ranks instead of leaving them in the broadcast
            import traceback

            box[0] = ("ERR", f"{type(e).__name__}: {e}\n{traceback.format_exc()[-1500:]}")
    if nranks > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    if isinstance(box[0], tuple):
        raise RuntimeError(f"setup on rank 0 failed: {box[0][1]}")
    d = box[0]
    with open(os.path.join(d, f"lp_{rank}.pkl"), "rb") as f:
        lp = pickle.load(f)
    os.remove(os.path.join(d, f"lp_{rank}.pkl"))
    if nranks > 1:
        dist.barrier(group=group)
    if rank == 0:
        shutil.rmtree(d, ignore_errors=True)
    return lp
