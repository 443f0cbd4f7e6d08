"""Dry run of bench.py's own-arm control flow on a box without a GPU: the CUDA context and the
torch.cuda calls are replaced by inert fakes, so this only guards the host-side Python of the
measurement contract (keys of the JSON line, workload plumbing) — no timing, no arithmetic."""
import argparse
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Info:
    solve_ms = 12.5
    outer_iterations = 7
    inner_iterations = 70
    mass_iterations = 3
    final_residual = 1e-9
    kernel_launches = 1234


class _FakeApi:
    def solve(self, h, rhs, x, info):
        return 0


class _FakeCtx:
    def __init__(self, cfg):
        self.config, self.api, self._h = cfg, _FakeApi(), None

    def augment_rhs(self, x):
        return np.asarray(x, dtype=np.float64).copy()

    def solve_dev(self, d_rhs, d_x):
        return _Info()

    def time_kernel(self, what, param=0, warmup=3, reps=20, flush_l2=True):
        return 0.1, 1.0e8, 1

    def nccl_unique_id(self):
        return bytes(128)


@pytest.mark.skipif(torch.cuda.is_available(), reason="dry run is for GPU-less boxes")
def test_own_arm_control_flow_emits_the_contract_keys(monkeypatch, capfd):
    import bench
    import fictitious_domain_al_preconditioners_b200 as pkg
    from fictitious_domain_al_preconditioners_b200 import partition as part

    monkeypatch.setattr(pkg, "ALContext", _FakeCtx)
    monkeypatch.setattr(part, "setup_local_context", lambda ctx, lp, uid=bytes(128): ctx)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    real_zeros = torch.zeros
    monkeypatch.setattr(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(bench, "cpu_sample", lambda prob, H, threads, outer_steps=2: (0.5, 2))
    lines = []
    monkeypatch.setattr(bench, "emit", lambda s: lines.append(s))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0})
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", workload="tiny", nel=0, no_cpu=False,
                              no_graphs=False, no_bsr=False, expected_outer=0)
    bench.run_ours(args, dict(bench.WORKLOADS["tiny"]), "tiny")
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["config"]["workload"] == "tiny" and d["config"]["block_size"] == 2 and d["n_gpus"] == 1
    assert d["roofline"]["bound"] == "hbm" and d["roofline"]["unit"] == "GB/s" and "frac" in d["roofline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * d["config"]["n_dofs"]
    assert d["gpu_launches"] == 2 * 1234 and d["dtype"] == "f64" and d["vs_baseline"] is None


def _two_rank_worker(rank, port, q):
    import unittest.mock as mock

    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import bench
    import fictitious_domain_al_preconditioners_b200 as pkg
    from fictitious_domain_al_preconditioners_b200 import partition as part

    real_init = dist.init_process_group
    real_zeros = torch.zeros
    real_tensor = torch.tensor
    lines = []
    try:
        with mock.patch.object(pkg, "ALContext", _FakeCtx), \
                mock.patch.object(part, "setup_local_context", lambda ctx, lp, uid=bytes(128): ctx), \
                mock.patch.object(torch.cuda, "set_device", lambda *a, **k: None), \
                mock.patch.object(torch.cuda, "synchronize", lambda *a, **k: None), \
                mock.patch.object(torch.Tensor, "cuda", lambda self, *a, **k: self), \
                mock.patch.object(torch.Tensor, "pin_memory", lambda self, *a, **k: self), \
                mock.patch.object(torch, "zeros", lambda *a, **k: real_zeros(*a, **{kk: v for kk, v in k.items() if kk != "device"})), \
                mock.patch.object(torch, "tensor", lambda *a, **k: real_tensor(*a, **{kk: v for kk, v in k.items() if kk != "device"})), \
                mock.patch.object(dist, "init_process_group", lambda backend, **k: real_init("gloo")), \
                mock.patch.object(bench, "emit", lambda s: lines.append(s)), \
                mock.patch.object(bench.ClockSampler, "start", lambda self: None), \
                mock.patch.object(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}):
            args = argparse.Namespace(gpus=2, steps=1, warmup=1, impl="ours", workload="tiny", nel=0, no_cpu=True,
                                      no_graphs=False, no_bsr=False, expected_outer=0)
            bench.run_ours(args, dict(bench.WORKLOADS["tiny"]), "tiny")
        q.put((rank, lines))
    except Exception:
        import traceback

        q.put((rank, "ERR " + traceback.format_exc()[-2500:]))


@pytest.mark.skipif(torch.cuda.is_available(), reason="dry run is for GPU-less boxes")
def test_two_rank_control_flow_rank0_setup_and_single_json_line():
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_two_rank_worker, args=(r, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert not isinstance(out[0], str), out[0]
    assert not isinstance(out[1], str), out[1]
    assert len(out[0]) == 1 and len(out[1]) == 0  # rank 0 alone prints
    d = json.loads(out[0][0])
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and "weak-scaled x2" in d["config"]["description"]
    assert d["config"]["n_dofs"] > 9605  # the global job grew with the rank count
    assert d["config"]["setup"].startswith("rank 0 builds")
