#!/bin/bash
# 1-GPU call: new tests (device CSR -> BSR conversion, Chebyshev mass solves with the measured defaults), the whole
# -m gpu suite, the bench lines the changes are about and — only if every test passed — the headline bench with the
# phase timers of fdal_finalize.
cd /root/repo || exit 1
mkdir -p gpurun_out
export FDAL_PARITY_TAG=1gpu_c
rm -f gpurun_out/parity_log_1gpu_c.jsonl
{
  echo "== gpu tests"
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 2>&1 | grep -v Warning | tail -30
  rc=${PIPESTATUS[0]}
  echo "pytest rc=$rc"
  echo "== elliptic cycle 6 / cycle 5 (defaults), configs[1]"
  timeout 400 python bench.py --workload elliptic --cycle 6 --steps 2 --warmup 1 --no-parity 2>gpurun_out/fin_elliptic6.err | tee gpurun_out/fin_elliptic6.json | cut -c1-160
  timeout 300 python bench.py --workload elliptic --cycle 5 --steps 2 --warmup 1 --no-parity 2>gpurun_out/fin_elliptic5.err | tee gpurun_out/fin_elliptic5.json | cut -c1-160
  FDAL_VERBOSE_SETUP=1 timeout 400 python bench.py --workload stokes2d_1M --steps 3 --warmup 2 2>gpurun_out/fin_s2d1M.err | tee gpurun_out/fin_s2d1M.json | cut -c1-160
  FDAL_HOST_BSR=1 FDAL_VERBOSE_SETUP=1 timeout 600 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-parity 2>gpurun_out/fin_s3d40_hostbsr.err | tee gpurun_out/fin_s3d40_hostbsr.json | cut -c1-160
  if [ "$rc" = "0" ]; then
    echo "== headline: 3-D Stokes IB nel=74, N=1 (default bench)"
    FDAL_VERBOSE_SETUP=1 timeout 1500 python bench.py --steps 3 --warmup 3 2>gpurun_out/fin_head_n1.err | tee gpurun_out/fin_head_n1.json | cut -c1-300
  fi
  grep -h "bench \|fdal_finalize\]" gpurun_out/fin_*.err | cut -c1-200
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv
} > gpurun_out/r2_final.log 2>&1
tail -150 gpurun_out/r2_final.log
