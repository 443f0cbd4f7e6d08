#!/bin/bash
# 1-GPU call: device CSR -> BSR conversion after the shared-memory opt-in fix; then the whole suite and, if green,
# the headline bench with the phase timers of fdal_finalize.
cd /root/repo || exit 1
mkdir -p gpurun_out
export FDAL_PARITY_TAG=1gpu_d
rm -f gpurun_out/parity_log_1gpu_d.jsonl
{
  echo "== device BSR tests"
  timeout 600 python -m pytest tests/test_gpu_bsr_build.py -q --maxfail=3 2>&1 | grep -v Warning | tail -40
  rc1=${PIPESTATUS[0]}
  echo "== 3-D nel=40, device BSR (compare gpurun_out/fin_s3d40_hostbsr.err)"
  FDAL_VERBOSE_SETUP=1 timeout 600 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-parity 2>gpurun_out/fin2_s3d40.err | tee gpurun_out/fin2_s3d40.json | cut -c1-160
  grep -h "bench \|fdal_finalize\]\|bsr_build" gpurun_out/fin2_s3d40.err | cut -c1-200
  if [ "$rc1" = "0" ]; then
    echo "== whole gpu suite"
    timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_bsr_build.py 2>&1 | grep -v Warning | tail -12
    rc2=${PIPESTATUS[0]}
    if [ "$rc2" = "0" ]; then
      echo "== headline: 3-D Stokes IB nel=74, N=1 (default bench)"
      FDAL_VERBOSE_SETUP=1 timeout 1500 python bench.py --steps 3 --warmup 3 2>gpurun_out/fin2_head_n1.err | tee gpurun_out/fin2_head_n1.json | cut -c1-300
      grep -h "bench \|fdal_finalize\]\|bsr_build" gpurun_out/fin2_head_n1.err | cut -c1-200
    fi
  fi
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv
} > gpurun_out/r2_final2.log 2>&1
tail -150 gpurun_out/r2_final2.log
