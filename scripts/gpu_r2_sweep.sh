#!/bin/bash
# Round-2 opening measurement (one gpurun call, 1 GPU, ~15 min): what round 1 left unmeasured.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/gpu_r2_sweep.sh'
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  echo "== BSR-2 tuning sweep on configs[1] (TPR x blocks-in-flight), CSR node-major for comparison"
  PROBE_CONFIGS="FDAL_BSR_TPR=2;FDAL_BSR_TPR=4;FDAL_BSR_TPR=8;FDAL_BSR_TPR=2,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=8,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=2;FDAL_NO_BSR=1,FDAL_TPR=8;FDAL_NO_BSR=1,FDAL_TPR=8,FDAL_SPMV_PF=1;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4,FDAL_SPMV_PF=1" \
    timeout 600 python scripts/kernel_probe.py stokes2d_1M 2>&1 | tail -12
  echo "== BSR-3 (3-D Stokes nel=32): blocks-in-flight"
  PROBE_CONFIGS="FDAL_BSR_TPR=16;FDAL_BSR_TPR=16,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=8,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4" \
    timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -6
  echo "== scalar CSR kernels with prefetch (Laplace r=10)"
  PROBE_CONFIGS="FDAL_TPR=2;FDAL_TPR=2,FDAL_SPMV_PF=1;FDAL_TPR=4,FDAL_SPMV_PF=1" \
    timeout 600 python scripts/kernel_probe.py laplace 2>&1 | tail -5
  echo "== default bench; with 4 blocks per lane in flight; with the dense W^-1 GEMV; with all three opt-ins"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2_bench_default.err | tee gpurun_out/r2_bench_default.json
  FDAL_BSR_UNROLL=4 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2_bench_unroll4.err | tee gpurun_out/r2_bench_unroll4.json
  FDAL_DENSE_WINV=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2_bench_dense_winv.err | tee gpurun_out/r2_bench_dense_winv.json
  FDAL_BSR_UNROLL=4 FDAL_DENSE_WINV=1 FDAL_SPMV_PF=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2_bench_all.err | tee gpurun_out/r2_bench_all.json
  echo "== gpu tests (incl. the non-strict xfail ones written after the last round-1 GPU run)"
  timeout 1200 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -60
} > gpurun_out/r2_sweep.log 2>&1
tail -100 gpurun_out/r2_sweep.log
