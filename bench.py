#!/usr/bin/env python
"""bench.py — outer FGMRES solve time / DoFs/s of the AL solve path on B200.

One "step" = one complete outer solve (FGMRES right-preconditioned by the
block-triangular AL preconditioner, inner PCG + AMG V-cycle) of the synthetic
refinement of the named parameter file.  Prints ONE JSON line (see the contract in
DESIGN.md "Measurement").

  python bench.py                       # N=1, default workload, few steps
  python bench.py --impl reference      # the CPU restatement (oracle) on host cores
  torchrun ... bench.py --gpus N ...    # one rank per GPU (row-partitioned solve)
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


METRIC = "outer_fgmres_solve_dofs_per_s"
UNIT = "DoF/s"

WORKLOADS = {
    # configs[1]: stokes_immersed_boundary 2D, parameters_stokes.prm as shipped, ~1M DoFs
    "stokes2d_1M": dict(kind="stokes", dim=2, nel=320, diagonal_mass=False,
                        label="stokes_immersed_boundary 2D parameters_stokes.prm (exact mass inverses), Q2^2-Q1 nel=320"),
    # configs[3] family: stokes_immersed_boundary 3D, parameters_stokes_3d.prm (diagonal W^-1, lumped-CG Mp^-1)
    "stokes3d": dict(kind="stokes", dim=3, nel=32, diagonal_mass=True,
                     label="stokes_immersed_boundary 3D parameters_stokes_3d.prm, Q2^3-Q1 nel=32"),
    "stokes2d_diag": dict(kind="stokes", dim=2, nel=320, diagonal_mass=True,
                          label="stokes_immersed_boundary 2D, diagonal mass, Q2^2-Q1 nel=320"),
    "laplace": dict(kind="laplace", r_bg=10, label="immersed_laplace 2D circle, Q1 r=10"),
    # configs[4] family: elasticity.prm (vector-valued elliptic interface, BSR-3 path)
    "elasticity": dict(kind="elasticity", dim=3, nel=64, label="elliptic_interface elasticity.prm, Q1^3 vector nel=64"),
    "tiny": dict(kind="stokes", dim=2, nel=32, diagonal_mass=True, label="tiny smoke workload"),
}


def build_problem(w):
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn

    if w["kind"] == "stokes":
        # node-major numbering: the velocity block and the finest AMG operator go to BSR
        prob = syn.stokes_immersed_boundary(dim=w["dim"], nel=w["nel"], diagonal_mass=w["diagonal_mass"],
                                            numbering="node")
    elif w["kind"] == "elasticity":
        prob = syn.elasticity_interface(nel_bg=w["nel"], nel_imm=max(2, w["nel"] // 4), diagonal_inverse=True)
    else:
        prob = syn.immersed_laplace(r_bg=w["r_bg"], diagonal_inverse=True)
    H = syn.build_hierarchies(prob)
    return prob, H


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


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_sample(prob, H, threads, outer_steps=2):
    """Time the oracle on host cores on a bounded sample: the first `outer_steps` outer
    iterations of the SAME solve, extrapolated with the full iteration count."""
    import copy

    from fictitious_domain_al_preconditioners_b200 import synthetic as syn
    from oracle import oracle

    cfg = copy.deepcopy(prob.config)
    cfg.outer.max_steps = outer_steps
    p2 = copy.copy(prob)
    p2.config = cfg
    ctx = syn.setup_context(oracle.OracleContext(cfg, threads=threads), p2, H, oracle=True)
    rhs = ctx.augment_rhs(prob.rhs) if prob.augment_rhs else prob.rhs
    t0 = time.perf_counter()
    _, info = ctx.solve(rhs, raise_on_failure=False)
    dt = time.perf_counter() - t0
    its = max(1, info.outer_iterations)
    ctx.close()
    return dt / its, its  # seconds per outer iteration


CPU_SAMPLE_MAX_DOFS = 1_500_000


def bounded_cpu_problem(w, prob, H):
    """The CPU arm must finish in minutes: beyond ~1.5 M DoFs the oracle is timed on a coarser
    refinement of the same parameter file (CPU DoF/s does not improve with size, so this
    does not flatter the GPU)."""
    if prob.n_dofs <= CPU_SAMPLE_MAX_DOFS or "nel" not in w:
        return prob, H, "the same solve"
    w2 = dict(w)
    dim = w.get("dim", 2)
    w2["nel"] = max(8, int(w["nel"] * (1.0e6 / prob.n_dofs) ** (1.0 / dim) / 2) * 2)
    sprob, sH = build_problem(w2)
    return sprob, sH, f"the same parameter file at nel={w2['nel']} ({sprob.n_dofs} DoFs)"


def estimate_dofs(w):
    """DoFs of a Stokes workload without building it (the reference arm only needs the count)."""
    if w.get("kind") != "stokes":
        return None
    dim, nel = w["dim"], w["nel"]
    if dim == 2:
        m = 2 * (2 ** int(round(np.log2(nel))) + 1)
    else:
        k = max(1, int(round(np.log2(nel))) - 3)
        m = 3 * (6 * 4**k + 2)
    return dim * (2 * nel + 1) ** dim + (nel + 1) ** dim + m


def run_reference(args, w, wname):
    """--impl reference: the reference's CPU path.  The reference binary cannot be
    built here (deal.II / Trilinos / UMFPACK absent), so this times the oracle port
    with all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle

    api = oracle.load()
    cores = max(1, min(api.get_max_threads(), os.cpu_count() or 1))
    world = max(1, args.gpus)
    if world > 1 and "nel" in w:  # the same weak-scaled job as our arm at this GPU count
        w = dict(w)
        w["nel"] = int(round(w["nel"] * world ** (1.0 / w["dim"]) / 2.0)) * 2
        w["label"] = w["label"].rsplit("nel=", 1)[0] + f"nel={w['nel']} (weak-scaled x{world})"
    n_est = estimate_dofs(w)
    if n_est is not None and n_est > CPU_SAMPLE_MAX_DOFS:
        # do not assemble the big problem just to count its unknowns
        w2 = dict(w)
        w2["nel"] = max(8, int(w["nel"] * (1.0e6 / n_est) ** (1.0 / w["dim"]) / 2) * 2)
        prob, H = build_problem(w2)
        n_full = n_est
        note = f"the same parameter file at nel={w2['nel']} ({prob.n_dofs} DoFs)"
    else:
        prob, H = build_problem(w)
        n_full = prob.n_dofs
        prob, H, note = bounded_cpu_problem(w, prob, H)
    sample_outer = 2
    vals = []
    for i in range(args.warmup + args.steps):
        per_it, _ = cpu_sample(prob, H, cores, sample_outer)
        if i >= args.warmup:
            vals.append(per_it)
    per_it = float(np.mean(vals))
    n_outer = args.expected_outer or estimate_outer(w)
    t_solve = per_it * n_outer
    value = prob.n_dofs / t_solve
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": n_full / value * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "description": w["label"], "n_dofs": n_full},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample_outer} outer FGMRES iterations of {note} on {cores} OpenMP threads, "
                                   f"extrapolated to {n_outer} outer iterations", "n_dofs_sample": prob.n_dofs},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


_OUTER_CACHE = {}


def estimate_outer(w):
    """Outer iteration count of the workload (mesh independent): taken from a coarse
    refinement of the same parameter file solved by the oracle."""
    key = json.dumps(w, sort_keys=True)
    if key in _OUTER_CACHE:
        return _OUTER_CACHE[key]
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn
    from oracle import oracle

    w2 = dict(w)
    if w["kind"] == "elasticity":
        w2["nel"] = 16
    elif w["kind"] == "stokes":
        w2["nel"] = 32 if w["dim"] == 2 else 8
    else:
        w2["r_bg"] = 6
    prob, H = build_problem(w2)
    ctx = syn.setup_context(oracle.OracleContext(prob.config), prob, H, oracle=True)
    rhs = ctx.augment_rhs(prob.rhs) if prob.augment_rhs else prob.rhs
    _, info = ctx.solve(rhs, raise_on_failure=False)
    _OUTER_CACHE[key] = max(1, info.outer_iterations)
    return _OUTER_CACHE[key]


def run_ours(args, w, wname):
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    # torchrun pins OMP_NUM_THREADS=1; the host-side setup helpers (OpenMP SpGEMM, BSR
    # conversion) should share the box's cores between the ranks instead
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // max(1, world_env)))
    if world_env > 1:
        # every rank builds the hierarchy itself: keep the setup bit-reproducible across ranks
        # (no GPU mat-vec in the eigenvalue estimate) so all ranks cut identical matrices
        os.environ["FDAL_DETERMINISTIC_SETUP"] = "1"
    import torch

    from fictitious_domain_al_preconditioners_b200 import ALContext
    from fictitious_domain_al_preconditioners_b200 import _binding as b
    from fictitious_domain_al_preconditioners_b200 import synthetic as syn

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist

        # a rank that fails leaves the others inside an NCCL call for ever: bound the damage
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
    if world > 1 and "nel" in w:
        # weak scaling: per-GPU work fixed, the global grid grows with the GPU count
        w = dict(w)
        w["nel"] = int(round(w["nel"] * world ** (1.0 / w["dim"]) / 2.0)) * 2
        w["label"] = w["label"].rsplit("nel=", 1)[0] + f"nel={w['nel']} (weak-scaled x{world})"
    from fictitious_domain_al_preconditioners_b200 import partition as part

    prob = H = None

    def build():
        p_, H_ = build_problem(w)
        meta = dict(n_dofs=int(p_.n_dofs), sizes=[int(x) for x in p_.sizes], nnz_A=int(p_.A.nnz),
                    amg_levels=H_[0].describe())
        return p_, H_, meta

    uid = [bytes(128)]
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
    else:
        prob, H, meta = build()
        lp = part.distribute_problem(prob, H, 0, 1)
        lp.rhs_local, lp.augment_rhs, lp.meta = lp.scatter(prob.rhs), bool(prob.augment_rhs), meta
    t_gen = time.perf_counter() - t0
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
    rhs = lp.rhs_local
    if lp.augment_rhs:
        rhs = ctx.augment_rhs(rhs)
    N = rhs.size
    n_dofs_global = lp.meta["n_dofs"]
    t_setup = time.perf_counter() - t0

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs already in HBM --------------------------------
    d_rhs = torch.from_numpy(rhs).cuda()
    d_x = torch.zeros(N, dtype=torch.float64, device="cuda")
    infos = []
    for _ in range(args.warmup):
        d_x.zero_()
        torch.cuda.synchronize()
        ctx.solve_dev(d_rhs.data_ptr(), d_x.data_ptr())
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        d_x.zero_()
        torch.cuda.synchronize()
        info = ctx.solve_dev(d_rhs.data_ptr(), d_x.data_ptr())
        dev_ms += info.solve_ms
        infos.append(info)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    ms_step = dev_ms / args.steps
    if world > 1:
        import torch.distributed as dist

        tt = torch.tensor([ms_step], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())
    x_dev = d_x.cpu().numpy()

    # ---- end-to-end arm: host buffers through the reference-facing C-ABI call ---------
    h_rhs = torch.from_numpy(rhs).pin_memory()
    h_x = torch.zeros(N, dtype=torch.float64).pin_memory()
    e2e_t = []
    for i in range(min(args.warmup, 1) + args.steps):
        h_x.zero_()
        barrier()
        t0 = time.perf_counter()
        info = b.SolveInfo()
        import ctypes as C

        st = ctx.api.solve(ctx._h, C.cast(h_rhs.data_ptr(), C.POINTER(C.c_double)),
                           C.cast(h_x.data_ptr(), C.POINTER(C.c_double)), C.byref(info))
        assert st == 0, st
        barrier()
        if i >= min(args.warmup, 1):
            e2e_t.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_t))
    if world > 1:
        import torch.distributed as dist

        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())

    # ---- per-kernel roofline, timed live with CUDA events on the library's stream ------
    peak, peak_src = measured_peak()
    kern = {}
    for name, what, param in (("cheb_fine", b.TIME_CHEB_FINE, 0), ("spmv_A", b.TIME_SPMV_A, 0),
                              ("aug_apply", b.TIME_AUG, 0), ("vcycle", b.TIME_VCYCLE, 0),
                              ("dot", b.TIME_DOT, 0), ("multidot16", b.TIME_MULTIDOT, 16), ("axpy", b.TIME_AXPY, 0)):
        try:
            ms, by, nl = ctx.time_kernel(what, param, warmup=3, reps=20, flush_l2=True)
            kern[name] = {"ms": ms, "alg_bytes": by, "GBps": by / ms * 1e-6, "frac": by / ms * 1e-6 / peak,
                          "launches": nl}
        except Exception as e:  # e.g. multidot without FGMRES basis
            kern[name] = {"error": str(e)}
    dom = kern.get("cheb_fine", {})
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        if wname in tj and world == 1 and not args.nel and not args.no_bsr:
            traffic = tj[wname]["traffic_bytes_per_launch"]
    except Exception:
        pass
    last = infos[-1]
    total_dofs = n_dofs_global  # the whole (weak-scaled) job, all ranks together
    if rank == 0:
        res = {
            "metric": METRIC, "value": total_dofs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "description": w["label"], "n_dofs": n_dofs_global, "blocks": lp.meta["sizes"],
                       "parallelism": f"row-partitioned x{world}" if world > 1 else "single GPU",
                       "nnz_A": lp.meta["nnz_A"], "amg_levels": lp.meta["amg_levels"],
                       "l2_policy": "working set (matrices + hierarchy) larger than L2; kernel timings flush L2 "
                                    "with a 256 MiB memset between launches",
                       "outer_iterations": int(last.outer_iterations), "inner_iterations": int(last.inner_iterations),
                       "mass_iterations": int(last.mass_iterations), "final_residual": last.final_residual,
                       "setup_s": {"generate+amg_host": t_gen, "upload+finalize": t_setup},
                       "wall_ms_per_step": wall / args.steps * 1e3, "graphs": bool(cfg.use_graphs),
                       "block_size": int(cfg.block_size),
                       "setup": "rank 0 builds and cuts the problem; shares via /dev/shm" if world > 1 else "in process"},
            "e2e": {"value": total_dofs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 16 * N, "d2h_bytes_per_step": 8 * N,
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(sum(i.kernel_launches for i in infos)),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "fused Chebyshev step on the finest AMG level (k_bsr_spmv<B,TPR,EpiCheb> / k_spmv<TPR,EpiCheb>)",
                         "achieved": dom.get("GBps"), "peak": peak, "unit": "GB/s",
                         "frac": dom.get("frac"), "traffic": traffic, "algorithmic_bytes": dom.get("alg_bytes"),
                         "frac_of_nominal_8000GBs": (dom.get("GBps") / 8000.0) if dom.get("GBps") else None,
                         "peak_source": peak_src},
            "kernels": kern,
        }
        if not args.no_cpu and world == 1:
            n_outer = int(last.outer_iterations)
            sprob, sH, note = bounded_cpu_problem(w, prob, H)
            per_it, _ = cpu_sample(sprob, sH, threads=1, outer_steps=2)
            res["cpu_baseline"] = {"value": sprob.n_dofs / (per_it * n_outer), "unit": UNIT, "cores": 1, "kind": "port",
                                   "sample": f"first 2 outer FGMRES iterations of {note} (oracle, 1 thread = how the "
                                             f"reference ships), extrapolated to {n_outer} outer iterations",
                                   "ms_per_step_sample_problem": per_it * n_outer * 1e3, "n_dofs_sample": sprob.n_dofs}
        emit(json.dumps(res))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FDAL_BENCH_WORKLOAD", "stokes2d_1M"))
    ap.add_argument("--nel", type=int, default=0, help="override the refinement of the workload")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-bsr", action="store_true")
    ap.add_argument("--expected-outer", type=int, default=0)
    args = ap.parse_args()
    capture_stdout()
    w = dict(WORKLOADS[args.workload])
    if args.nel:
        w["nel"] = args.nel
        w["label"] = w["label"].rsplit("nel=", 1)[0] + f"nel={args.nel}"
    if args.impl == "reference":
        run_reference(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
