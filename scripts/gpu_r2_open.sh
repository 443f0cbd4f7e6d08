#!/bin/bash
# Round-2 opening call (1 GPU): box facts, the sweep round 1 left unmeasured, a 3-D nel=64 run with setup timings.
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  echo "== box"; nproc; free -g; df -h /dev/shm /tmp | cat; nvidia-smi --query-gpu=name,memory.total --format=csv; ulimit -a | head -20
  echo "== BSR-2 tuning sweep on configs[1]"
  PROBE_CONFIGS="FDAL_BSR_TPR=2;FDAL_BSR_TPR=4;FDAL_BSR_TPR=8;FDAL_BSR_TPR=2,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=8,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=2;FDAL_NO_BSR=1,FDAL_TPR=8;FDAL_NO_BSR=1,FDAL_TPR=8,FDAL_SPMV_PF=1;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4,FDAL_SPMV_PF=1;FDAL_BSR_TPR=4,FDAL_DENSE_WINV=1" \
    timeout 600 python scripts/kernel_probe.py stokes2d_1M 2>&1 | tail -14
  echo "== BSR-3 (3-D Stokes nel=32)"
  PROBE_CONFIGS="FDAL_BSR_TPR=16;FDAL_BSR_TPR=16,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=8,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=4;FDAL_BSR_TPR=8;FDAL_BSR_TPR=16,FDAL_SPMV_PF=1" \
    timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -8
  echo "== scalar CSR kernels with prefetch (Laplace r=10)"
  PROBE_CONFIGS="FDAL_TPR=2;FDAL_TPR=2,FDAL_SPMV_PF=1;FDAL_TPR=4,FDAL_SPMV_PF=1" \
    timeout 600 python scripts/kernel_probe.py laplace 2>&1 | tail -5
  echo "== default 2-D bench with the dense W^-1 GEMV"
  FDAL_DENSE_WINV=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/r2_bench_dense_winv.err | tee gpurun_out/r2_bench_dense_winv.json
  echo "== 3-D nel=64 with setup timings"
  FDAL_VERBOSE_SETUP=1 timeout 1200 python bench.py --workload stokes3d --nel 64 --steps 2 --warmup 1 --no-cpu 2>gpurun_out/r2_s3d64.err | tee gpurun_out/r2_s3d64.json
  tail -40 gpurun_out/r2_s3d64.err
  echo "== gpu tests"
  timeout 1200 python -m pytest tests -m gpu -q -rxX 2>&1 | tail -70
} > gpurun_out/r2_open.log 2>&1
tail -150 gpurun_out/r2_open.log
