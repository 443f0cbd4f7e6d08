#!/bin/bash
# 2-GPU call: the peer-channel multi-GPU path against the oracle (+ the single-GPU parity suite on the changed kernels)
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  nvidia-smi topo -m | head -12
  echo "== single-GPU parity (kernels changed)"
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bsr.py -x -q 2>&1 | tail -8
  echo "== multi-GPU"
  timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q -s 2>&1 | grep -v Warning | tail -40
} > gpurun_out/r2_p2p.log 2>&1
tail -60 gpurun_out/r2_p2p.log
