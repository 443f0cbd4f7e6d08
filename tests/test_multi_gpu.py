"""Row-partitioned solve on 2 (and 4) GPUs against the serial CPU oracle, in both communication
modes: peer channels (cudaIpc-mapped halo / all-reduce / all-gather buffers written by the library's
own kernels, the default) and the NCCL send/recv + all-reduce fallback.  Needs >= 2 CUDA devices:
skipped otherwise (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, nranks, port, case, mode, q):
    import threading

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["FDAL_COMM"] = "nccl" if mode == "nccl" else "p2p"
    if mode == "p2p_deep":  # partition every AMG level but the coarsest
        os.environ["FDAL_REP_ROWS"] = "0"
    # a rank stuck in a spinning kernel would otherwise outlive the test
    wd = threading.Timer(180, lambda: os._exit(7))
    wd.daemon = True
    wd.start()
    dist.init_process_group("gloo", rank=rank, world_size=nranks)
    try:
        from fictitious_domain_al_preconditioners_b200 import ALContext
        from fictitious_domain_al_preconditioners_b200 import partition as part
        from fictitious_domain_al_preconditioners_b200 import synthetic as syn

        from tests import problems as P

        torch.cuda.set_device(rank)
        prob, H = P.get(case)
        prob.config.device = rank
        lp = part.distribute_problem(prob, H, rank, nranks)
        ctx = ALContext(prob.config)
        uid = [ctx.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        part.setup_local_context(ctx, lp, uid[0])
        res = {"comm_mode": ctx.api.comm_mode(ctx._h), "rep_from": lp.amg[0].rep_from, "levels": len(lp.amg[0].levels)}
        # operator / preconditioner applications on scattered random vectors
        X = P.rand(prob.n_dofs, 5)
        y_loc = ctx.apply_system(lp.scatter(X))
        u = P.rand(prob.n_dofs, 10)
        v_loc, its = ctx.apply_prec(lp.scatter(u))
        r0 = P.rand(prob.sizes[0], 7)
        r0_full = np.concatenate([r0, np.zeros(prob.n_dofs - r0.size)])
        z_loc = ctx.apply_amg(lp.scatter(r0_full)[: lp.sizes_local[0]])
        # full solve: the rhs is augmented on the device, per rank
        rhs_loc = lp.scatter(prob.rhs)
        if prob.augment_rhs:
            rhs_loc = ctx.augment_rhs(rhs_loc)
        x_loc, info = ctx.solve(rhs_loc)
        gathered = [None] * nranks
        dist.all_gather_object(gathered, (y_loc, v_loc, z_loc, x_loc, rhs_loc))
        if rank == 0:
            from oracle import oracle

            ora = syn.setup_context(oracle.OracleContext(prob.config), prob, H, oracle=True)
            Y = lp.gather([g[0] for g in gathered])
            V = lp.gather([g[1] for g in gathered])
            zpad = [np.concatenate([g[2], np.zeros(g[0].size - g[2].size)]) for g in gathered]
            Z = lp.gather(zpad)[: prob.sizes[0]]
            Xs = lp.gather([g[3] for g in gathered])
            RHS = lp.gather([g[4] for g in gathered])
            res["system"] = P.relerr(Y, ora.apply_system(X))
            vo, ito = ora.apply_prec(u)
            res["prec"] = P.relerr(V, vo)
            res["prec_its"] = (tuple(its), tuple(ito))
            res["amg"] = P.relerr(Z, ora.apply_amg(r0))
            rhs_o = P.rhs_of(ora, prob)
            res["rhs"] = P.relerr(RHS, rhs_o)
            xo, io = ora.solve(rhs_o)
            res["solve"] = P.relerr(Xs, xo)
            res["outer"] = (info.outer_iterations, io.outer_iterations)
            res["inner"] = (info.inner_iterations, io.inner_iterations)
            q.put(res)
        ctx.close()
    finally:
        dist.destroy_process_group()


CASES = [("laplace_diag", 2, "p2p"), ("stokes2d_diag", 2, "p2p"), ("stokes3d_diag", 2, "p2p"),
         ("stokes2d_exact", 2, "p2p"), ("elliptic_modified_diag", 2, "p2p"), ("elliptic_ideal", 2, "p2p"),
         ("stokes2d_diag", 2, "p2p_deep"), ("stokes3d_diag", 2, "p2p_deep"),
         ("stokes2d_diag", 2, "nccl"), ("stokes3d_diag", 2, "nccl"),
         ("stokes3d_diag", 4, "p2p"), ("stokes2d_diag", 4, "p2p_deep"), ("laplace_diag", 4, "p2p")]


@pytest.mark.parametrize("case,nranks,mode", CASES)
def test_multi_gpu_solve_matches_oracle(case, nranks, mode):
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, nranks, port, case, mode, q)) for r in range(nranks)]
    for p in procs:
        p.start()
    try:
        res = q.get(timeout=200)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0
    print(case, nranks, mode, res)
    from tests import parity_log

    parity_log.record(f"multi_gpu[{case},{nranks},{mode}]", res)
    assert res["comm_mode"] == (1 if mode == "nccl" else 2)
    assert res["system"] < 1e-12
    assert res["amg"] < 1e-12
    assert res["rhs"] < 1e-12
    assert res["prec_its"][0] == res["prec_its"][1]
    assert res["prec"] < 1e-9
    assert abs(res["outer"][0] - res["outer"][1]) <= 1
    if res["outer"][0] == res["outer"][1]:
        # identical inner trajectories reproduce the solution to solver accuracy; where one inner CG stopped an
        # iteration apart (block-CG of the ideal variant: see the self-sensitivity note in test_gpu_parity.py) the
        # two runs are different, equally valid, approximate preconditioners
        assert res["solve"] < (1e-8 if res["inner"][0] == res["inner"][1] else 1e-5)
