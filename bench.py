#!/usr/bin/env python
"""bench.py — outer FGMRES solve time / DoFs/s of the AL solve path on B200.

One "step" = one complete outer solve (FGMRES right-preconditioned by the block-triangular AL
preconditioner, inner PCG + AMG V-cycle) of the synthetic refinement of the named parameter file.
Prints ONE JSON line (contract: DESIGN.md "Measurement").

  python bench.py                       # N=1, headline workload (3-D Stokes IB, 10.35 M DoFs)
  python bench.py --impl reference      # the CPU restatement (oracle): ONE COMPLETE solve, all host cores
  torchrun ... bench.py --gpus N ...    # the SAME problem row-partitioned over N GPUs (strong scaling)

Every timed step goes through the reference-facing C-ABI call `fdal_solve` with pinned HOST buffers
(H2D of right-hand side and initial guess, D2H of the solution inside the timed wall-clock region =
`e2e`); the CUDA events inside the same call bracket the device-resident part (inputs already in
HBM) = `value`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_RESULT_FD = None


def capture_stdout():
    """Native libraries write to fd 1 ("NCCL version ..." at communicator creation): route
    everything except the result line to stderr so stdout carries exactly ONE JSON line."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: str):
    sys.stdout.flush()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (line + "\n").encode())


def log(msg):
    sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


T_START = time.perf_counter()
# the driver gives every bench step a fixed wall-clock limit (870 s in the scaling run): both arms watch this
# deadline and shorten the run (fewer timed solves, reported as such) rather than be killed without a line
DEADLINE_S = float(os.environ.get("FDAL_BENCH_DEADLINE_S", "780"))


def elapsed():
    return time.perf_counter() - T_START


METRIC = "outer_fgmres_solve_dofs_per_s"
UNIT = "DoF/s"
DEFAULT_WORKLOAD = "stokes3d_10M"

WORKLOADS = {
    # configs[3] (headline, north_star): stokes_immersed_boundary 3D, parameters_stokes_3d.prm (diagonal W^-1,
    # lumped-CG Mp^-1), Q2^3-Q1 on 74^3 cells = 10.35 M DoFs; row-partitioned over 1/2/4/8 GPUs
    "stokes3d_10M": dict(kind="stokes", dim=3, nel=74, diagonal_mass=True,
                         label="stokes_immersed_boundary 3D parameters_stokes_3d.prm, Q2^3-Q1 nel=74"),
    "stokes3d": dict(kind="stokes", dim=3, nel=32, diagonal_mass=True,
                     label="stokes_immersed_boundary 3D parameters_stokes_3d.prm, Q2^3-Q1 nel=32"),
    # configs[1]: stokes_immersed_boundary 2D, parameters_stokes.prm as shipped (exact mass inverses), ~1M DoFs
    "stokes2d_1M": dict(kind="stokes", dim=2, nel=320, diagonal_mass=False,
                        label="stokes_immersed_boundary 2D parameters_stokes.prm (exact mass inverses), Q2^2-Q1 nel=320"),
    "stokes2d_diag": dict(kind="stokes", dim=2, nel=320, diagonal_mass=True,
                          label="stokes_immersed_boundary 2D, diagonal mass, Q2^2-Q1 nel=320"),
    # configs[0]: immersed_laplace 2D circle as shipped (Circle_parameters_f1_g1 family: matrix-free
    # K + gamma Ct (M^-1 M^-1) C with exact mass inverses, immersed_laplace.cc:875-884)
    "laplace": dict(kind="laplace", r_bg=10, diagonal_inverse=False,
                    label="immersed_laplace 2D circle, exact (M^-1)^2 as shipped, Q1 r=10"),
    "laplace_diag": dict(kind="laplace", r_bg=10, diagonal_inverse=True,
                         label="immersed_laplace 2D circle, diagonal W^-1, Q1 r=10"),
    # configs[2]: elliptic_interface co-dim 0, parameters_ideal.prm (modified AL, h-scaled mass, exact M^-1),
    # --beta2 sweeps the coefficient jump, --cycle the refinement (cycle 7 = 4.2 M background DoFs)
    "elliptic": dict(kind="elliptic", cycle=6, beta2=1e3,
                     label="elliptic_interface parameters_ideal.prm (modified AL), cycle=6"),
    # configs[4] family: elasticity.prm (vector-valued elliptic interface, BSR-3 path)
    "elasticity": dict(kind="elasticity", dim=3, nel=64, label="elliptic_interface elasticity.prm, Q1^3 vector nel=64"),
    "tiny": dict(kind="stokes", dim=2, nel=32, diagonal_mass=True, label="tiny smoke workload, Q2^2-Q1 nel=32"),
    "tiny3d": dict(kind="stokes", dim=3, nel=8, diagonal_mass=True, label="tiny 3-D smoke workload, Q2^3-Q1 nel=8"),
}


def build_problem(w):
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn

    if w["kind"] == "stokes":
        # node-major numbering: the velocity block and the finest AMG operator go to BSR
        prob = syn.stokes_immersed_boundary(dim=w["dim"], nel=w["nel"], diagonal_mass=w["diagonal_mass"],
                                            numbering="node")
    elif w["kind"] == "elasticity":
        prob = syn.elasticity_interface(nel_bg=w["nel"], nel_imm=max(2, w["nel"] // 4), diagonal_inverse=True)
    elif w["kind"] == "elliptic":
        prob = syn.elliptic_interface(cycle=w["cycle"], beta2=w.get("beta2", 1e3))
    else:
        prob = syn.immersed_laplace(r_bg=w["r_bg"], diagonal_inverse=w.get("diagonal_inverse", False))
    H = syn.build_hierarchies(prob)
    return prob, H


def mass_forms(ctx):
    """How the exact mass inverses (W^-1 = M^-1 / M^-2, Mp^-1) are applied on the device (fdal_mass_solver_info)."""
    out = {}
    for which, name in ((0, "M"), (1, "Mp")):
        try:
            d = ctx.mass_solver_info(which)
            if d["form"] != "none":
                out[name] = d
        except Exception:
            pass
    return out


def config_of(wname, w, n_dofs, world, scaling):
    """The keys BOTH arms print (identical for the same job)."""
    return {"workload": wname, "description": w["label"], "n_dofs": int(n_dofs), "n_gpus_job": int(world),
            "scaling_mode": scaling}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak(sustained=False):
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def weak_scaled(w, world):
    if world <= 1 or "nel" not in w:
        return w
    w = dict(w)
    w["nel"] = int(round(w["nel"] * world ** (1.0 / w["dim"]) / 2.0)) * 2
    w["label"] = w["label"].rsplit("nel=", 1)[0] + f"nel={w['nel']} (weak-scaled x{world})"
    return w


# --------------------------------------------------------------------------------------- reference arm
def run_reference(args, w, wname):
    """--impl reference: the reference's CPU path.  The reference binary cannot be built here
    (deal.II / Trilinos / UMFPACK absent), so this times the oracle port with all host
    threads.  A step is ONE COMPLETE outer solve of the benched problem — measured, nothing
    extrapolated.  The run stops early once FDAL_REF_BUDGET_S (default 600 s) of solving has been
    spent (always at least one complete solve); `steps` / `warmup` report what was actually run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn
    from oracle import oracle

    api = oracle.load()
    cores = max(1, min(api.get_max_threads(), os.cpu_count() or 1))
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    world = max(1, args.gpus)
    if args.scaling == "weak":
        w = weak_scaled(w, world)
    t0 = time.perf_counter()
    prob, H = build_problem(w)
    t_setup = time.perf_counter() - t0
    log(f"reference arm: problem + hierarchy in {t_setup:.1f} s, {prob.n_dofs} DoFs")
    budget = float(os.environ.get("FDAL_REF_BUDGET_S", "600"))
    times, infos, spent, n_warm = [], [], 0.0, 0
    want_warm = args.warmup
    extrapolated = None
    ctx = syn.setup_context(oracle.OracleContext(prob.config, threads=cores), prob, H, oracle=True)
    rhs = ctx.augment_rhs(prob.rhs) if prob.augment_rhs else prob.rhs
    if prob.n_dofs > int(float(os.environ.get("FDAL_REF_PROBE_MIN_DOFS", "2e6"))):
        # a complete solve of a problem this size takes minutes.  Go / no-go against the step's wall-clock limit
        # from one V-cycle + one augmented apply (the work of one inner PCG iteration) times the inner
        # iterations such a solve takes (FDAL_EXPECTED_INNER, default 400: the CUDA arm needs 361 on the headline problem)
        r = np.random.default_rng(0).uniform(-1, 1, prob.sizes[0])
        t0 = time.perf_counter()
        ctx.apply_amg(r)
        ctx.apply_aug(r)
        t_it = time.perf_counter() - t0
        n_inner = int(os.environ.get("FDAL_EXPECTED_INNER", "400"))
        predicted = 1.08 * t_it * n_inner
        log(f"reference arm: one inner iteration's work takes {t_it:.2f} s; a complete solve is predicted at {predicted:.0f} s, "
            f"{DEADLINE_S - elapsed():.0f} s left before the deadline")
        if predicted * 1.15 + 20 > DEADLINE_S - elapsed():
            # last resort: time the first two outer iterations (own context with max_steps = 2) and extrapolate
            import copy

            ctx.close()
            ctx = None
            cfg2 = copy.deepcopy(prob.config)
            cfg2.outer.max_steps = 2
            p2 = copy.copy(prob)
            p2.config = cfg2
            pctx = syn.setup_context(oracle.OracleContext(cfg2, threads=cores), p2, H, oracle=True)
            t0 = time.perf_counter()
            _, pinfo = pctx.solve(rhs, raise_on_failure=False)
            t_probe = time.perf_counter() - t0
            pctx.close()
            n_outer_gpu = int(os.environ.get("FDAL_EXPECTED_OUTER", "0")) or 12
            extrapolated = dict(t_probe=t_probe, outer_probe=int(pinfo.outer_iterations), inner_probe=int(pinfo.inner_iterations),
                                n_outer=n_outer_gpu, predicted=t_probe / max(1, pinfo.outer_iterations) * n_outer_gpu)
    while extrapolated is None and len(times) < max(1, args.steps):
        t0 = time.perf_counter()
        _, info = ctx.solve(rhs, raise_on_failure=False)
        dt = time.perf_counter() - t0
        spent += dt
        log(f"reference arm: complete solve in {dt:.2f} s ({info.outer_iterations} outer / {info.inner_iterations} inner)")
        # a warm-up solve is only affordable when a solve is short compared with the budget
        if n_warm < want_warm and spent + dt * (1 + len(times)) < budget * 0.5 and elapsed() + 2 * dt < DEADLINE_S:
            n_warm += 1
            continue
        times.append(dt)
        infos.append(info)
        if spent + dt > budget or elapsed() + dt > DEADLINE_S:
            break
    if ctx is not None:
        ctx.close()
    if extrapolated is not None:
        # last resort (stated in the line): the complete solve does not fit the step's wall-clock limit
        t_solve = extrapolated["predicted"]
        times = []

        class _I:
            outer_iterations, inner_iterations = extrapolated["n_outer"], 0
            final_residual, status = float("nan"), -1

        info = _I()
    else:
        t_solve = float(np.mean(times))
        info = infos[-1]
    value = prob.n_dofs / t_solve
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": max(1, len(times)), "warmup": n_warm, "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": t_solve * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wname, w, prob.n_dofs, world, args.scaling),
        "solve": {"outer_iterations": int(info.outer_iterations), "inner_iterations": int(info.inner_iterations),
                  "final_residual": info.final_residual, "status": int(info.status), "setup_s": t_setup,
                  "solve_s_each": times},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": (f"{len(times)} COMPLETE outer solve(s) of the benched problem ({prob.n_dofs} DoFs) by the "
                                    f"oracle port on {cores} OpenMP threads after {n_warm} warm-up solve(s); measured, not "
                                    f"extrapolated; run bounded by a {budget:.0f} s solve budget") if extrapolated is None else
                                   (f"EXTRAPOLATED (the complete solve was predicted at {extrapolated['predicted']:.0f} s and did not "
                                    f"fit the {DEADLINE_S:.0f} s step limit): first {extrapolated['outer_probe']} outer iterations "
                                    f"({extrapolated['inner_probe']} inner) of the benched problem took {extrapolated['t_probe']:.1f} s "
                                    f"on {cores} OpenMP threads, scaled to {extrapolated['n_outer']} outer iterations"),
                         "extrapolated": extrapolated is not None},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# --------------------------------------------------------------------------------------- CPU-side checks
def scipy_apply_system(prob, x):
    """AA x of the Stokes / Laplace block system straight from the scipy blocks (independent of the
    oracle and of the CUDA library): the N>1 parity check of the partitioned operator."""
    from fictitious_domain_al_preconditioners_b200 import _binding as b

    kind = prob.config.kind
    n, m = prob.Ct.shape
    g = prob.config.gamma
    x0 = x[:n]
    cx = prob.Ct.T @ x0
    if prob.config.winv_mode != b.WINV_DIAG or prob.config.aug_explicit:
        return None
    t = g * prob.winv_diag * cx
    if kind == b.KIND_LAPLACE:
        return np.concatenate([prob.A @ x0 + prob.Ct @ (t + x[n:]), cx])
    if kind in (b.KIND_STOKES, b.KIND_STOKES_DIAG_MINRES) and not prob.config.grad_div_in_operator:
        n_p = prob.Bt.shape[1]
        x1, x2 = x[n:n + n_p], x[n + n_p:]
        return np.concatenate([prob.A @ x0 + prob.Ct @ (t + x2) + prob.Bt @ x1, prob.Bt.T @ x0, cx])
    return None


def oracle_checks(prob, H, ctx_gpu, lp, x_solution, rhs, n_inner, threads_all):
    """N=1: parity of one block-system apply and one V-cycle against the oracle at the benched size,
    the true residual of the GPU solution, and the bounded CPU sample (one inner PCG iteration's
    work — V-cycle + augmented apply — on 1 thread and on all threads)."""
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn
    from oracle import oracle

    t0 = time.perf_counter()
    ora = syn.setup_context(oracle.OracleContext(prob.config, threads=threads_all), prob, H, oracle=True)
    log(f"oracle context for the parity object / CPU sample: {time.perf_counter() - t0:.1f} s")
    rng = np.random.default_rng(0)
    N, n0 = prob.n_dofs, prob.sizes[0]
    x = rng.uniform(-1, 1, N)
    r = rng.uniform(-1, 1, n0)

    def rel(a, ref):
        return float(np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-300))

    par = {}
    yo = ora.apply_system(x)
    par["apply_system_vs_oracle"] = rel(lp.gather([ctx_gpu.apply_system(lp.scatter(x))]), yo)
    xs = np.concatenate([r, np.zeros(N - n0)])
    zg = ctx_gpu.apply_amg(lp.scatter(xs)[:n0])
    zo = ora.apply_amg(r)
    par["apply_amg_vs_oracle"] = rel(lp.gather([np.concatenate([zg, np.zeros(N - n0)])])[:n0], zo)
    ao = ora.apply_aug(r)
    ag = ctx_gpu.apply_aug(lp.scatter(xs)[:n0])
    par["apply_aug_vs_oracle"] = rel(lp.gather([np.concatenate([ag, np.zeros(N - n0)])])[:n0], ao)
    xg = lp.gather([x_solution])
    rhs_g = lp.gather([rhs])
    par["true_residual_rel"] = rel(ora.apply_system(xg), rhs_g)
    par["n_dofs"] = int(N)
    par["tolerances"] = {"apply": 1e-12, "solution": 1e-10}
    # bounded CPU sample
    sample = {}
    for thr in (1, threads_all):
        oracle.load().set_num_threads(thr)
        reps = 1 if thr == 1 else 3
        t0 = time.perf_counter()
        for _ in range(reps):
            ora.apply_amg(r)
            ora.apply_aug(r)
        sample[thr] = (time.perf_counter() - t0) / reps
    ora.close()
    return par, sample


# --------------------------------------------------------------------------------------- our arm
def run_ours(args, w, wname):
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun pins OMP_NUM_THREADS=1; the host-side setup helpers (OpenMP SpGEMM, BSR
    # conversion) should share the box's cores between the ranks instead
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, world_env)))
    import torch

    from fictitious_domain_al_preconditioners_b200 import ALContext
    from fictitious_domain_al_preconditioners_b200 import _binding as b

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        # a rank that fails leaves the others inside a spinning kernel for ever: bound the damage
        limit = float(os.environ.get("FDAL_BENCH_WATCHDOG_S", "1500"))

        def _bail():
            sys.stderr.write(f"[bench] rank {rank}: watchdog fired after {limit:.0f} s, aborting\n")
            sys.stderr.flush()
            os._exit(3)

        wd = threading.Timer(limit, _bail)
        wd.daemon = True
        wd.start()
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    t0 = time.perf_counter()
    if args.scaling == "weak":
        w = weak_scaled(w, world)
    from fictitious_domain_al_preconditioners_b200 import partition as part

    prob = H = None
    keep = {}

    def build():
        p_, H_ = build_problem(w)
        meta = dict(n_dofs=int(p_.n_dofs), sizes=[int(x) for x in p_.sizes], nnz_A=int(p_.A.nnz),
                    amg_levels=H_[0].describe())
        keep["prob"] = p_  # rank 0 keeps the global blocks for the N>1 parity check
        return p_, H_, meta

    uid = [bytes(128)]
    gloo = None
    if world > 1:
        import torch.distributed as dist

        # setup on rank 0 only (one copy of the global problem in host memory, one hierarchy
        # for every rank's halo plan); the shares travel through shared-memory files
        from fictitious_domain_al_preconditioners_b200 import amg_setup

        gloo = dist.new_group(backend="gloo")
        ncpu = os.cpu_count() or 1
        amg_setup.set_host_threads(ncpu if rank == 0 else 1)  # rank 0 sets up with the whole box
        lp = part.share_local_problems(build, rank, world, gloo)
        amg_setup.set_host_threads(max(1, ncpu // world))  # then every rank gets its share
        prob = keep.get("prob")
    else:
        prob, H, meta = build()
        lp = part.distribute_problem(prob, H, 0, 1)
        lp.rhs_local, lp.augment_rhs, lp.meta = lp.scatter(prob.rhs), bool(prob.augment_rhs), meta
    t_gen = time.perf_counter() - t0
    torch.cuda.empty_cache()  # the generators' torch scratch must not sit in the caching allocator beside the library
    log(f"rank {rank}: problem + hierarchy + partition in {t_gen:.1f} s")
    cfg = lp.config
    cfg.device = local_rank
    cfg.use_graphs = not args.no_graphs
    if not args.no_bsr:
        cfg.block_size = lp.block_size
    t0 = time.perf_counter()
    ctx = ALContext(cfg)
    if world > 1:
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, group=gloo)
    part.setup_local_context(ctx, lp, uid[0])
    comm_mode = {0: "single GPU", 1: "nccl send/recv + all-reduce", 2: "peer channels (cudaIpc, in-kernel NVLink stores)"}[
        ctx.api.comm_mode(ctx._h)]
    rhs = lp.rhs_local
    if lp.augment_rhs:
        rhs = ctx.augment_rhs(rhs)
    N = rhs.size
    n_dofs_global = lp.meta["n_dofs"]
    t_setup = time.perf_counter() - t0
    log(f"rank {rank}: upload + finalize in {t_setup:.1f} s ({comm_mode})")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return float(v)
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- the timed steps: fdal_solve with pinned HOST buffers; CUDA events inside bracket the device part
    import ctypes as C

    h_rhs = torch.from_numpy(rhs).pin_memory()
    h_x = torch.zeros(N, dtype=torch.float64).pin_memory()
    p_rhs = C.cast(h_rhs.data_ptr(), C.POINTER(C.c_double))
    p_x = C.cast(h_x.data_ptr(), C.POINTER(C.c_double))

    def one_step():
        h_x.zero_()
        info = b.SolveInfo()
        st = ctx.api.solve(ctx._h, p_rhs, p_x, C.byref(info))
        if st != 0:
            raise RuntimeError(f"fdal_solve status {st}: {ctx.api.last_error(ctx._h)}")
        return info

    # the first solve also tells how many more fit before the deadline (same decision on every rank)
    n_warm, n_steps = args.warmup, args.steps
    t0 = time.perf_counter()
    first = one_step()
    t_first = allmax(time.perf_counter() - t0)
    reserve = 45.0 + (0.6 * (t_gen + t_setup) if (world == 1 and not args.no_parity) else 20.0)
    fit = int(max(1.0, (DEADLINE_S - allmax(elapsed()) - reserve) / max(t_first, 1e-3)))
    if fit < n_warm - 1 + n_steps:
        n_warm = max(1, min(n_warm, fit // 4 + 1))
        n_steps = max(1, min(n_steps, fit - (n_warm - 1)))
        log(f"rank {rank}: a solve takes {t_first:.1f} s, {fit} fit before the deadline: {n_warm} warm-up + {n_steps} timed steps "
            f"instead of {args.warmup} + {args.steps}")
    for _ in range(max(0, n_warm - 1)):
        one_step()
    if n_warm == 0:
        n_warm = 1  # the probing solve above was one
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_region = time.perf_counter()
    infos, walls = [], []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        infos.append(one_step())
        barrier()
        walls.append(time.perf_counter() - t0)
    barrier()
    wall_total = time.perf_counter() - t_region
    clocks = sampler.stop()
    ms_step = allmax(float(np.mean([i.solve_ms for i in infos])))
    e2e_s = allmax(float(np.mean(walls)))
    x_loc = h_x.numpy().copy()
    last = infos[-1]
    log(f"rank {rank}: {n_steps} steps, {ms_step:.1f} ms/solve device, {e2e_s*1e3:.1f} ms end to end, "
        f"{last.outer_iterations} outer / {last.inner_iterations} inner")

    # ---- per-kernel roofline, timed live with CUDA events on the library's stream ------
    peak, peak_src = measured_peak()
    kern = {}
    for name, what, param in (("cheb_fine", b.TIME_CHEB_FINE, 0), ("spmv_A", b.TIME_SPMV_A, 0),
                              ("aug_apply", b.TIME_AUG, 0), ("vcycle", b.TIME_VCYCLE, 0),
                              ("dot", b.TIME_DOT, 0), ("multidot16", b.TIME_MULTIDOT, 16), ("axpy", b.TIME_AXPY, 0)):
        try:
            ms, by, nl = ctx.time_kernel(what, param, warmup=3, reps=20, flush_l2=True)
            ms = allmax(ms)
            kern[name] = {"ms": ms, "alg_bytes": by, "GBps": by / ms * 1e-6, "frac": by / ms * 1e-6 / peak,
                          "launches": nl}
        except Exception as e:  # e.g. multidot without FGMRES basis
            kern[name] = {"error": str(e)}
    dom = kern.get("cheb_fine", {})
    traffic, traffic_note = None, None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if wname in tj and world == 1 and not args.nel and not args.no_bsr:
            traffic = tj[wname]["traffic_bytes_per_launch"]
        elif w["kind"] == "stokes" and w.get("dim") == 3 and not args.no_bsr and "stokes3d_nel40" in tj:
            # no capture of this refinement (ncu replays every kernel ~40 times: too long at 10 M DoFs): `traffic` stays
            # null, the measured traffic / algorithmic-bytes ratio of the same kernel on the same operator at nel=40 rides along
            traffic_note = {"captured_on": "stokes3d nel=40", "dram_bytes_over_algorithmic_bytes": tj["stokes3d_nel40"]["ratio"],
                            "source": tj["stokes3d_nel40"]["source"]}
    except Exception:
        pass

    # ---- parity object: correctness travels with every line (also in the scaling runs) ----
    parity, cpu_sample = None, None
    gathered = None
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((x_loc, rhs), gathered, dst=0, group=gloo)
    skip_oracle = world == 1 and elapsed() + 0.6 * (t_gen + t_setup) > DEADLINE_S
    if rank == 0 and not args.no_parity:
        try:
            if skip_oracle:
                parity = {"skipped": f"{elapsed():.0f} s into a {DEADLINE_S:.0f} s step: no time left for the oracle context"}
            elif world == 1:
                parity, cpu_sample = oracle_checks(prob, H, ctx, lp, x_loc, rhs, int(last.inner_iterations),
                                                   max(1, os.cpu_count() or 1))
            else:
                xg = lp.gather([g[0] for g in gathered])
                rg = lp.gather([g[1] for g in gathered])
                parity = {"n_dofs": int(n_dofs_global)}
                res = scipy_apply_system(prob, xg)
                if res is not None:
                    parity["true_residual_rel_scipy"] = float(np.linalg.norm(res - rg) / np.linalg.norm(rg))
                parity["note"] = ("solution of the partitioned solve gathered on rank 0; AA x - b evaluated with scipy on the "
                                  "global blocks (independent of the CUDA library and of the oracle)")
        except Exception as e:
            parity = {"error": f"{type(e).__name__}: {e}"}
    if world > 1:
        # the partitioned operator against scipy on a random vector (halo exchange, all-reduce of C x)
        xr = np.random.default_rng(1).uniform(-1, 1, n_dofs_global)
        y_loc = ctx.apply_system(lp.scatter(xr))
        gy = [None] * world if rank == 0 else None
        dist.gather_object(y_loc, gy, dst=0, group=gloo)
        if rank == 0 and parity is not None and "error" not in parity:
            ref = scipy_apply_system(prob, xr)
            if ref is not None:
                parity["apply_system_vs_scipy"] = float(np.linalg.norm(lp.gather(gy) - ref) / np.linalg.norm(ref))

    if rank == 0:
        res = {
            "metric": METRIC, "value": n_dofs_global / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": n_steps, "warmup": n_warm, "requested": {"steps": args.steps, "warmup": args.warmup},
            "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wname, w, n_dofs_global, world, args.scaling),
            "solve": {"outer_iterations": int(last.outer_iterations), "inner_iterations": int(last.inner_iterations),
                      "mass_iterations": int(last.mass_iterations), "final_residual": last.final_residual,
                      "ms_per_inner_iteration": ms_step / max(1, int(last.inner_iterations)),
                      "blocks": lp.meta["sizes"], "nnz_A": lp.meta["nnz_A"], "amg_levels": lp.meta["amg_levels"],
                      "parallelism": f"row-partitioned x{world}, {comm_mode}" if world > 1 else "single GPU",
                      "l2_policy": "working set (matrices + hierarchy) larger than L2; kernel timings flush L2 "
                                   "with a 256 MiB memset between launches",
                      "setup_s": {"generate+amg_host": t_gen, "upload+finalize": t_setup},
                      "timed_region_wall_s": wall_total, "graphs": bool(cfg.use_graphs),
                      "block_size": int(cfg.block_size),
                      "exact_mass_inverses": mass_forms(ctx),
                      "setup": "rank 0 builds and cuts the problem; shares via /dev/shm" if world > 1 else "in process"},
            "e2e": {"value": n_dofs_global / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 16 * N, "d2h_bytes_per_step": 8 * N,
                    "ms_per_step": e2e_s * 1e3,
                    "how": "wall clock around fdal_solve (pinned host rhs + initial guess in, solution out), max over ranks"},
            "gpu_launches": int(sum(i.kernel_launches for i in infos)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fused Chebyshev step on the finest AMG level (k_bsr_spmv<B,TPR,EpiCheb> / k_spmv<TPR,EpiCheb>)",
                         "achieved": dom.get("GBps"), "peak": peak, "unit": "GB/s",
                         "frac": dom.get("frac"), "traffic": traffic, "traffic_note": traffic_note,
                         "algorithmic_bytes": dom.get("alg_bytes"),
                         "frac_of_nominal_8000GBs": (dom.get("GBps") / 8000.0) if dom.get("GBps") else None,
                         "peak_source": peak_src},
            "kernels": kern,
            "parity": parity,
        }
        if cpu_sample is not None:
            n_inner = max(1, int(last.inner_iterations))
            t1 = cpu_sample[1] * n_inner
            tall = cpu_sample[max(cpu_sample)] * n_inner
            res["cpu_baseline"] = {
                "value": n_dofs_global / t1, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"one inner PCG iteration's dominant work (one AMG V-cycle + one augmented apply of the benched "
                          f"{n_dofs_global}-DoF problem) by the oracle on 1 thread (how the reference ships): "
                          f"{cpu_sample[1]:.2f} s, times the {n_inner} inner iterations of the solve; the complete CPU solve "
                          f"is what `--impl reference` measures",
                "all_cores": {"cores": max(cpu_sample), "value": n_dofs_global / tall, "s_per_inner_iteration": cpu_sample[max(cpu_sample)]},
                "s_per_inner_iteration": cpu_sample[1]}
        emit(json.dumps(res))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FDAL_BENCH_WORKLOAD", DEFAULT_WORKLOAD))
    ap.add_argument("--nel", type=int, default=0, help="override the refinement of the workload")
    ap.add_argument("--cycle", type=int, default=0, help="elliptic workload: refinement cycle")
    ap.add_argument("--beta2", type=float, default=0.0, help="elliptic workload: coefficient of the inclusion")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the same problem on every GPU count; weak: the grid grows with the GPU count")
    ap.add_argument("--no-cpu", "--no-parity", dest="no_parity", action="store_true",
                    help="skip the oracle parity object and the CPU sample")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-bsr", action="store_true")
    args = ap.parse_args()
    capture_stdout()
    w = dict(WORKLOADS[args.workload])
    if args.nel:
        w["nel"] = args.nel
        w["label"] = w["label"].rsplit("nel=", 1)[0] + f"nel={args.nel}"
    if args.cycle and "cycle" in w:
        w["cycle"] = args.cycle
        w["label"] = w["label"].rsplit("cycle=", 1)[0] + f"cycle={args.cycle}"
    if args.beta2 and w["kind"] == "elliptic":
        w["beta2"] = args.beta2
        w["label"] += f", beta2={args.beta2:g}"
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
