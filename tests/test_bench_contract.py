"""bench.py contract checks that do not need a GPU: the reference arm prints one JSON line with
the keys the driver reads, and the workload table / weak-scaling rule are consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_valid_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "outer_fgmres_solve_dofs_per_s" and d["unit"] == "DoF/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f64"
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    # a step of the reference arm is a COMPLETE measured solve; both arms print the same config keys
    assert d["steps"] == 1 and "COMPLETE" in d["cpu_baseline"]["sample"] and d["solve"]["status"] == 0
    assert set(d["config"]) == {"workload", "description", "n_dofs", "n_gpus_job", "scaling_mode"}
    assert abs(d["ms_per_step"] * 1e-3 * d["value"] - d["config"]["n_dofs"]) < 1e-6 * d["config"]["n_dofs"]


def test_workload_table_names_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench

    # the default is the headline config: >= 10 M-DoF 3-D Stokes IB (BASELINE configs[3])
    w = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert w["dim"] == 3 and 3 * (2 * w["nel"] + 1) ** 3 + (w["nel"] + 1) ** 3 >= 10_000_000
    assert bench.WORKLOADS["laplace"]["diagonal_inverse"] is False  # exact mass inverses as shipped
    assert bench.WORKLOADS["elliptic"]["kind"] == "elliptic"
    assert "stokes2d_1M" in bench.WORKLOADS and bench.WORKLOADS["stokes2d_1M"]["diagonal_mass"] is False
    assert bench.WORKLOADS["stokes3d"]["dim"] == 3 and bench.WORKLOADS["stokes3d"]["diagonal_mass"] is True
    assert bench.WORKLOADS["elasticity"]["kind"] == "elasticity"
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "stokes_immersed_boundary 2D" in base["configs"][1]
