"""Row-partitioned solve on 2 GPUs (NCCL halo exchange + all-reduce) against the serial
CPU oracle.  Needs >= 2 CUDA devices: skipped otherwise (run with `gpurun --gpus 2`)."""
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


def _worker(rank, nranks, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
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
        res = {}
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


@pytest.mark.parametrize("case", ["laplace_diag", "stokes2d_diag", "stokes3d_diag", "stokes2d_exact", "elliptic_modified_diag"])
def test_two_gpu_solve_matches_oracle(case):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print(case, res)
    assert res["system"] < 1e-12
    assert res["amg"] < 1e-12
    assert res["rhs"] < 1e-12
    assert res["prec_its"][0] == res["prec_its"][1]
    assert res["prec"] < 1e-9
    assert abs(res["outer"][0] - res["outer"][1]) <= 1
    if res["outer"][0] == res["outer"][1]:
        assert res["solve"] < 1e-8
