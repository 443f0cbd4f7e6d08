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


class _FakeApi:
    def solve(self, h, rhs, x, info_ref):
        i = info_ref._obj
        i.solve_ms, i.outer_iterations, i.inner_iterations, i.mass_iterations = 12.5, 7, 70, 3
        i.final_residual, i.kernel_launches = 1e-9, 1234
        return 0

    def comm_mode(self, h):
        return 0

    def last_error(self, h):
        return b""


class _FakeCtx:
    def __init__(self, cfg):
        self.config, self.api, self._h = cfg, _FakeApi(), None

    def augment_rhs(self, x):
        return np.asarray(x, dtype=np.float64).copy()

    def apply_system(self, x):
        return np.asarray(x, dtype=np.float64).copy()

    def time_kernel(self, what, param=0, warmup=3, reps=20, flush_l2=True):
        return 0.1, 1.0e8, 1

    def nccl_unique_id(self):
        return bytes(128)

    def close(self):
        pass


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
    monkeypatch.setattr(bench, "oracle_checks", lambda *a, **k: ({"apply_system_vs_oracle": 1e-15}, {1: 0.5, 8: 0.1}))
    lines = []
    monkeypatch.setattr(bench, "emit", lambda s: lines.append(s))
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0})
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", workload="tiny", nel=0, no_parity=False,
                              no_graphs=False, no_bsr=False, scaling="strong")
    bench.run_ours(args, dict(bench.WORKLOADS["tiny"]), "tiny")
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline",
              "parity", "solve"):
        assert k in d, k
    assert d["config"]["workload"] == "tiny" and d["solve"]["block_size"] == 2 and d["n_gpus"] == 1
    assert set(d["config"]) == {"workload", "description", "n_dofs", "n_gpus_job", "scaling_mode"}  # both arms print these
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] == "port"
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
            args = argparse.Namespace(gpus=2, steps=1, warmup=1, impl="ours", workload="tiny", nel=0, no_parity=False,
                                      no_graphs=False, no_bsr=False, scaling=os.environ.get("TEST_SCALING", "strong"))
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
    # strong scaling: the SAME problem on every GPU count (BASELINE configs[3])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and "weak-scaled" not in d["config"]["description"]
    assert d["config"]["n_dofs"] == 9605 and d["config"]["n_gpus_job"] == 2
    assert d["solve"]["setup"].startswith("rank 0 builds")
    # the fake context's apply_system is the identity, so this only checks that the N>1 parity object is produced
    assert "apply_system_vs_scipy" in d["parity"] and "true_residual_rel_scipy" in d["parity"]
