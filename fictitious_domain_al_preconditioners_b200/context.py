"""Host-side handle on one ``fdal_ctx`` (see ``include/fdal.h``).

``ALContext`` is a thin, typed wrapper: numpy / scipy in, numpy out.  It does
no arithmetic itself — every ``apply_*`` / ``solve`` call goes through the C
ABI into the CUDA library, which must be present (``lib.load()`` raises
otherwise; there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _binding as b


class FdalError(RuntimeError):
    """Non-zero status from the C ABI.  ``status`` carries the FDAL_ERR_* code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"fdal status {status}: {message}")
        self.status = status


class NoConvergence(FdalError):
    """Mirror of dealii::SolverControl::NoConvergence (inner or outer solver)."""


@dataclass
class SolverControl:
    """deal.II SolverControl / ReductionControl / IterationNumberControl."""

    max_steps: int = 100
    tol: float = 1e-10
    reduce: float = 0.0
    type: int = b.CONTROL_SOLVER

    def c(self) -> b.Control:
        return b.Control(self.type, self.max_steps, self.tol, self.reduce)


def ReductionControl(max_steps=100, tol=1e-10, reduce=1e-2):
    return SolverControl(max_steps, tol, reduce, b.CONTROL_REDUCTION)


def IterationNumberControl(max_steps=100, tol=1e-12):
    return SolverControl(max_steps, tol, 0.0, b.CONTROL_ITERATION_NUMBER)


@dataclass
class ALConfig:
    kind: int = b.KIND_LAPLACE
    restart: int = 30
    gamma: float = 10.0
    gamma2: float = 0.0
    gamma_grad_div: float = 0.0
    winv_mode: int = b.WINV_DIAG
    mp_inv_mode: int = b.MPINV_CG_LUMPED
    aug_explicit: bool = False
    grad_div_in_operator: bool = False
    inner_prec: int = b.PREC_AMG
    device: int = 0
    use_graphs: bool = True
    exact_mass_max_its: int = 0
    block_size: int = 1
    outer: SolverControl = field(default_factory=lambda: ReductionControl(1000, 1e-10, 1e-12))
    inner: SolverControl = field(default_factory=lambda: SolverControl(100, 1e-2))
    mass: SolverControl = field(default_factory=lambda: SolverControl(100, 1e-6))

    def c(self) -> b.Config:
        return b.Config(
            self.kind,
            self.restart,
            self.gamma,
            self.gamma2,
            self.gamma_grad_div,
            self.winv_mode,
            self.mp_inv_mode,
            int(self.aug_explicit),
            int(self.grad_div_in_operator),
            self.inner_prec,
            self.device,
            int(self.use_graphs),
            self.exact_mass_max_its,
            self.block_size,
            0,
            self.outer.c(),
            self.inner.c(),
            self.mass.c(),
        )


class ALContext:
    """One solver context: matrices + hierarchy on the device, vmult-style calls."""

    def __init__(self, config: ALConfig, api: b.Api | None = None):
        if api is None:
            from .lib import load

            api = load()
        self.api = api
        self.config = config
        self._h = C.c_void_p()
        cfg = config.c()
        st = api.create(C.byref(self._h), C.byref(cfg))
        if st != b.OK:
            raise FdalError(st, "fdal_create failed")
        self._finalized = False
        self.sizes = None
        self.matrices = set()
        self.rank, self.nranks = 0, 1

    # -- lifetime ----------------------------------------------------------
    def close(self):
        if self._h:
            self.api.destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        msg = self.api.last_error(self._h)
        return msg.decode() if msg else ""

    def _check(self, st: int):
        if st == b.OK:
            return
        msg = self.api.last_error(self._h)
        msg = msg.decode() if msg else ""
        if st in (b.ERR_INNER_NO_CONVERGENCE, b.ERR_OUTER_NO_CONVERGENCE, b.ERR_MASS_NO_CONVERGENCE):
            raise NoConvergence(st, msg)
        raise FdalError(st, msg)

    # -- setup -------------------------------------------------------------
    def set_csr(self, matrix_id: int, A):
        rp, ci, v = b.csr_arrays(A)
        self.matrices.add(matrix_id)
        self._check(
            self.api.set_csr(
                self._h,
                matrix_id,
                A.shape[0],
                A.shape[1],
                v.size,
                rp.ctypes.data_as(C.POINTER(C.c_int64)),
                ci.ctypes.data_as(C.POINTER(C.c_int32)),
                b.dptr(v),
            )
        )

    def set_diag(self, diag_id: int, d):
        d = b.as_f64(d)
        self._check(self.api.set_diag(self._h, diag_id, d.size, b.dptr(d)))

    def set_amg(self, which: int, hierarchy):
        """``hierarchy``: ``amg_setup.Hierarchy`` (levels with A, P, R, inv_diag, lambda_max)."""
        levels = hierarchy.levels
        for l, L in enumerate(levels[:-1]):
            Av, ka = b.csr_view(L.A)
            Pv, kp = b.csr_view(L.P)
            Rv, kr = b.csr_view(L.R) if L.R is not None else (None, None)
            idg = b.as_f64(L.inv_diag) if L.inv_diag is not None else None
            self._check(
                self.api.amg_set_level(
                    self._h,
                    which,
                    l,
                    C.byref(Av),
                    C.byref(Pv),
                    C.byref(Rv) if Rv is not None else None,
                    b.dptr(idg) if idg is not None else None,
                    float(L.lambda_max),
                    int(hierarchy.cheb_degree),
                    float(hierarchy.eig_ratio),
                )
            )
        Av, ka = b.csr_view(levels[-1].A)
        self._check(self.api.amg_set_coarse(self._h, which, len(levels) - 1, C.byref(Av)))

    # -- multi-GPU (one process per GPU) --------------------------------------------
    def nccl_unique_id(self) -> bytes:
        buf = (C.c_char * 128)()
        self._check(self.api.nccl_unique_id(buf))
        return bytes(buf.raw)

    def comm_init(self, uid: bytes, rank: int, nranks: int):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._check(self.api.comm_init(self._h, buf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def set_halo(self, matrix_id: int, plan, level: int = 0, which: int = 0):
        i32 = C.POINTER(C.c_int32)
        sc = np.ascontiguousarray(plan.send_counts, dtype=np.int32)
        si = np.ascontiguousarray(plan.send_idx, dtype=np.int32)
        rc = np.ascontiguousarray(plan.recv_counts, dtype=np.int32)
        if si.size == 0:
            si = np.zeros(1, dtype=np.int32)
        self._check(self.api.set_halo(self._h, matrix_id, level, which, plan.n_owned, plan.n_halo,
                                      sc.ctypes.data_as(i32), si.ctypes.data_as(i32), rc.ctypes.data_as(i32)))

    def set_amg_local(self, which: int, LH):
        """``LH``: ``partition.LocalHierarchy`` (this rank's rows of every level)."""
        for l, L in enumerate(LH.levels):
            Av, ka = b.csr_view(L.A.local)
            Pv, kp = b.csr_view(L.P.local)
            Rv, kr = b.csr_view(L.R.local)
            idg = b.as_f64(L.inv_diag) if L.inv_diag is not None else None
            self._check(self.api.amg_set_level(
                self._h, which, l, C.byref(Av), C.byref(Pv), C.byref(Rv),
                b.dptr(idg) if idg is not None else None, float(L.lambda_max), int(LH.cheb_degree),
                float(LH.eig_ratio)))
            for mid, dc in ((b.MAT_AMG_A, L.A), (b.MAT_AMG_P, L.P), (b.MAT_AMG_R, L.R)):
                if dc.plan is not None:
                    self.set_halo(mid, dc.plan, level=l, which=which)
        Av, ka = b.csr_view(LH.coarse_A)
        nl = len(LH.levels)
        self._check(self.api.amg_set_coarse(self._h, which, nl, C.byref(Av)))
        if self.nranks > 1:
            rep_from = LH.rep_from if LH.rep_from >= 0 else nl
            self._check(self.api.amg_set_replicated_from(self._h, which, int(rep_from)))
            self._check(self.api.amg_set_coarse_range(self._h, which, int(LH.coarse_off[self.rank]),
                                                      int(LH.coarse_off[self.rank + 1])))

    def finalize(self):
        self._check(self.api.finalize(self._h))
        sizes = (C.c_int64 * 3)()
        nb = C.c_int()
        self._check(self.api.block_sizes(self._h, C.byref(sizes), C.byref(nb)))
        self.sizes = tuple(int(s) for s in sizes[: nb.value])
        self.N = sum(self.sizes)
        self._finalized = True

    # -- vmult-style calls ----------------------------------------------------
    def _out(self, n):
        return np.empty(n, dtype=np.float64)

    def spmv(self, matrix_id: int, x, transpose=False, n_out=None):
        x = b.as_f64(x)
        y = self._out(n_out)
        self._check(self.api.spmv(self._h, matrix_id, int(transpose), b.dptr(x), b.dptr(y)))
        return y

    def apply_aug(self, x, which=b.AMG_A11):
        x = b.as_f64(x)
        y = self._out(x.size)
        self._check(self.api.apply_aug(self._h, which, b.dptr(x), b.dptr(y)))
        return y

    def apply_system(self, x):
        x = b.as_f64(x)
        assert x.size == self.N
        y = self._out(self.N)
        self._check(self.api.apply_system(self._h, b.dptr(x), b.dptr(y)))
        return y

    def apply_winv(self, x):
        x = b.as_f64(x)
        y = self._out(x.size)
        self._check(self.api.apply_winv(self._h, b.dptr(x), b.dptr(y)))
        return y

    def apply_mp_inv(self, x):
        x = b.as_f64(x)
        y = self._out(x.size)
        its = C.c_int()
        self._check(self.api.apply_mp_inv(self._h, b.dptr(x), b.dptr(y), C.byref(its)))
        return y, its.value

    def apply_amg(self, r, which=b.AMG_A11):
        r = b.as_f64(r)
        z = self._out(r.size)
        self._check(self.api.apply_amg(self._h, which, b.dptr(r), b.dptr(z)))
        return z

    def apply_aug_inv(self, rhs, which=b.AMG_A11):
        rhs = b.as_f64(rhs)
        x = self._out(rhs.size)
        its = C.c_int()
        self._check(self.api.apply_aug_inv(self._h, which, b.dptr(rhs), b.dptr(x), C.byref(its)))
        return x, its.value

    def apply_prec(self, u):
        u = b.as_f64(u)
        assert u.size == self.N
        v = self._out(self.N)
        its = (C.c_int * 2)()
        self._check(self.api.apply_prec(self._h, b.dptr(u), b.dptr(v), C.byref(its)))
        return v, (its[0], its[1])

    def augment_rhs(self, rhs):
        rhs = b.as_f64(rhs).copy()
        self._check(self.api.augment_rhs(self._h, b.dptr(rhs)))
        return rhs

    def solve(self, rhs, x0=None, raise_on_failure=True):
        rhs = b.as_f64(rhs)
        assert rhs.size == self.N
        x = np.zeros(self.N) if x0 is None else b.as_f64(x0).copy()
        info = b.SolveInfo()
        st = self.api.solve(self._h, b.dptr(rhs), b.dptr(x), C.byref(info))
        if st != b.OK and raise_on_failure:
            self._check(st)
        return x, info

    # -- device-pointer variants (CUDA library only) --------------------------------
    def solve_dev(self, d_rhs: int, d_x: int, raise_on_failure=True):
        info = b.SolveInfo()
        st = self.api.solve_dev(self._h, d_rhs, d_x, C.byref(info))
        if st != b.OK and raise_on_failure:
            self._check(st)
        return info

    def time_kernel(self, what: int, param=0, warmup=3, reps=20, flush_l2=True):
        ms = C.c_double()
        by = C.c_double()
        nl = C.c_int64()
        self._check(
            self.api.time_kernel(
                self._h, what, param, warmup, reps, int(flush_l2), C.byref(ms), C.byref(by), C.byref(nl)
            )
        )
        return ms.value, by.value, nl.value

    def bsr_conversions(self):
        """(on_device, on_host): where the CSR -> BSR conversions of ``fdal_finalize`` ran (CUDA library only)."""
        dev, host = C.c_int32(), C.c_int32()
        self._check(self.api.bsr_conversions(self._h, C.byref(dev), C.byref(host)))
        return dev.value, host.value

    MASS_FORMS = {0: "none", 1: "pcg_kernels", 2: "pcg_one_cta", 3: "dense", 4: "cheb_kernels", 5: "cheb_persistent"}

    def mass_solver_info(self, which: int = 0) -> dict:
        """How the exact inverse of the multiplier (which=0) / pressure (which=1) mass matrix is applied
        (``fdal_mass_solver_info``; CUDA library only)."""
        form, its = C.c_int32(), C.c_int32()
        lo, hi, res = C.c_double(), C.c_double(), C.c_double()
        self._check(self.api.mass_solver_info(self._h, which, C.byref(form), C.byref(its), C.byref(lo), C.byref(hi),
                                              C.byref(res)))
        return dict(form=self.MASS_FORMS.get(form.value, str(form.value)), iterations=its.value,
                    interval=(lo.value, hi.value), verified_residual=res.value)


def csr_to_bsr(A, block_size: int, max_fill: float = 1.35, device: int = 0, api: b.Api | None = None):
    """CSR -> BSR on the device (``fdal_csr_to_bsr``: the routine ``fdal_finalize`` uses for the dim-blocked
    matrices).  Returns (brow_ptr, bcol, bval[nblk, b, b]) or None when the matrix is not blocked."""
    import scipy.sparse as sp

    if api is None:
        from .lib import load

        api = load()
    A = sp.csr_matrix(A)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    v = np.ascontiguousarray(A.data, dtype=np.float64)
    n = A.shape[0]
    brp = np.zeros(n // block_size + 1, dtype=np.int32)
    p64, p32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    args = (device, n, A.nnz, rp.ctypes.data_as(p64), ci.ctypes.data_as(p32), b.dptr(v), block_size, max_fill,
            brp.ctypes.data_as(p32))
    nblk = api.csr_to_bsr(*args, 0, None, None)
    if nblk == -1000:
        return None
    if nblk < 0:
        raise FdalError(int(-nblk), "fdal_csr_to_bsr failed")
    bcj = np.zeros(max(nblk, 1), dtype=np.int32)
    bv = np.zeros(max(nblk, 1) * block_size * block_size, dtype=np.float64)
    got = api.csr_to_bsr(*args, nblk, bcj.ctypes.data_as(p32), b.dptr(bv))
    assert got == nblk
    return brp, bcj[:nblk], bv[: nblk * block_size * block_size].reshape(nblk, block_size, block_size)


def assemble_al_term(A, point_dofs, point_phi, weight, device: int = 0, api: b.Api | None = None):
    """``A + sum_q weight[q] phi_q phi_q^T`` on the device (``fdal_assemble_al_term``): the
    operator-form AL term of immersed_laplace.cc:659-702 scattered into the CSR values of ``A`` (scipy CSR;
    its pattern must already hold the couplings).  ``point_dofs`` / ``point_phi``: (n_points, dofs_per_cell)
    dof indices (negative = skip) and shape-function values of the background cell of every immersed
    quadrature point; ``weight``: gamma * JxW.  Returns a new scipy CSR matrix with the same pattern."""
    import scipy.sparse as sp

    if api is None:
        from .lib import load

        api = load()
    A = sp.csr_matrix(A)
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    v = np.array(A.data, dtype=np.float64, copy=True)
    dofs = np.ascontiguousarray(point_dofs, dtype=np.int32)
    phi = np.ascontiguousarray(point_phi, dtype=np.float64)
    w = np.ascontiguousarray(weight, dtype=np.float64)
    assert dofs.shape == phi.shape and dofs.ndim == 2 and w.size == dofs.shape[0]
    missing = C.c_int64(0)
    st = api.assemble_al_term(device, A.shape[0], rp.ctypes.data_as(C.POINTER(C.c_int64)),
                              ci.ctypes.data_as(C.POINTER(C.c_int32)), b.dptr(v), dofs.shape[0], dofs.shape[1],
                              dofs.ctypes.data_as(C.POINTER(C.c_int32)), b.dptr(phi), b.dptr(w), C.byref(missing))
    if st != b.OK:
        raise FdalError(st, f"fdal_assemble_al_term failed ({missing.value} entries missing from the sparsity pattern)")
    return sp.csr_matrix((v, A.indices.copy(), A.indptr.copy()), shape=A.shape)
