#!/bin/bash
# 1-GPU call: full -m gpu suite (parity log), kernel probes after the kernel restructuring, headline N=1 bench at 10.35 M DoFs
cd /root/repo || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_log.jsonl
{
  echo "== gpu tests"
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -12
  echo "== kernel probes"
  PROBE_CONFIGS="FDAL_BSR_TPR=4;FDAL_BSR_TPR=4,FDAL_NO_CHUNK_FLAGS=1;FDAL_BSR_TPR=4,FDAL_MERGE=1" timeout 600 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -4
  PROBE_CONFIGS="FDAL_BSR_TPR=16;FDAL_BSR_TPR=16,FDAL_NO_CHUNK_FLAGS=1;FDAL_BSR_TPR=16,FDAL_MERGE=1" timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -4
  echo "== headline: 3-D Stokes IB nel=74, N=1"
  FDAL_VERBOSE_SETUP=1 timeout 2400 python bench.py --steps 3 --warmup 1 2>gpurun_out/r2_head_n1.err | tee gpurun_out/r2_head_n1.json | cut -c1-400
  grep -E "bench |setup\]|Error|error|Traceback" gpurun_out/r2_head_n1.err | tail -50
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv
} > gpurun_out/r2_head1.log 2>&1
tail -90 gpurun_out/r2_head1.log
