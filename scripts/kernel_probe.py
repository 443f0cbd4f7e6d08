"""Times the SpMV-family kernels of one workload under different tuning knobs
(measurement helper; prints a table).  usage: kernel_probe.py <workload> [nel]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fictitious_domain_al_preconditioners_b200 import ALContext, _binding as b, partition as part, synthetic as syn  # noqa: E402

wname = sys.argv[1] if len(sys.argv) > 1 else "stokes2d_diag"
w = dict(bench.WORKLOADS[wname])
if len(sys.argv) > 2:
    w["nel"] = int(sys.argv[2])
prob, H = bench.build_problem(w)
peak, _ = bench.measured_peak()
configs = [c.split(",") for c in os.environ.get("PROBE_CONFIGS", "").split(";") if c] or [
    ["FDAL_SPMV=classic"],
    ["FDAL_SPMV=classic", "FDAL_UNROLL=4"],
    ["FDAL_SPMV=classic", "FDAL_UNROLL=4", "FDAL_TPR=16"],
    ["FDAL_SPMV=classic", "FDAL_UNROLL=4", "FDAL_TPR=8"],
    ["FDAL_SPMV=classic", "FDAL_UNROLL=4", "FDAL_TPR=4"],
    ["FDAL_SPMV=stream", "FDAL_STREAM_CTAS=3"],
    ["FDAL_SPMV=stream", "FDAL_STREAM_CTAS=2"],
]
KN = ("FDAL_SPMV", "FDAL_UNROLL", "FDAL_TPR", "FDAL_STREAM_CTAS", "FDAL_NO_BSR", "FDAL_BSR_TPR", "FDAL_BSR_AOS", "FDAL_BSR_UNROLL", "FDAL_SPMV_PF", "FDAL_DENSE_WINV")
lp = part.distribute_problem(prob, H, 0, 1)
prob.config.block_size = lp.block_size
print(f"workload {wname} N={prob.n_dofs} nnz(A)={prob.A.nnz} levels={H[0].describe()}")
for cfg in configs:
    for k in KN:
        os.environ.pop(k, None)
    for kv in cfg:
        k, v = kv.split("=")
        os.environ[k] = v
    ctx = part.setup_local_context(ALContext(prob.config), lp)
    row = {}
    for name, what in (("spmv_A", b.TIME_SPMV_A), ("cheb_fine", b.TIME_CHEB_FINE), ("aug", b.TIME_AUG),
                       ("vcycle", b.TIME_VCYCLE)):
        ms, by, nl = ctx.time_kernel(what, 0, warmup=3, reps=20, flush_l2=True)
        row[name] = f"{ms*1e3:7.1f}us {by/ms*1e-6/peak*100:5.1f}%"
    ctx.close()
    print(" ".join(cfg).ljust(60), json.dumps(row))
