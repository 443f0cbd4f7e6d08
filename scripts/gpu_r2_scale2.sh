#!/bin/bash
# 2-GPU call: remaining multi-GPU parity cases, then 1- vs 2-GPU timings (2-D weak-scaled, 3-D strong)
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  echo "== multi-GPU parity"
  timeout 1500 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | grep -v Warning | tail -15
  echo "== 2-D diag, N=1"
  timeout 600 python bench.py --workload stokes2d_diag --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_s2d_n1.err | tee gpurun_out/r2_s2d_n1.json | cut -c1-300
  echo "== 2-D diag, N=2 weak (peer channels)"
  timeout 900 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --workload stokes2d_diag --scaling weak --steps 3 --warmup 2 2>gpurun_out/r2_s2d_n2.err | tee gpurun_out/r2_s2d_n2.json | cut -c1-300
  echo "== 2-D diag, N=2 weak (NCCL fallback)"
  FDAL_COMM=nccl timeout 900 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --workload stokes2d_diag --scaling weak --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_s2d_n2_nccl.err | tee gpurun_out/r2_s2d_n2_nccl.json | cut -c1-300
  echo "== 3-D nel=40, N=1"
  timeout 900 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-parity 2>gpurun_out/r2_s3d40_n1.err | tee gpurun_out/r2_s3d40_n1.json | cut -c1-300
  echo "== 3-D nel=40, N=2 strong"
  timeout 900 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --workload stokes3d --nel 40 --steps 2 --warmup 1 2>gpurun_out/r2_s3d40_n2.err | tee gpurun_out/r2_s3d40_n2.json | cut -c1-300
  for f in gpurun_out/r2_s2d_n2.err gpurun_out/r2_s3d40_n2.err; do echo "-- $f"; grep "bench \|Error\|error" $f | tail -12; done
} > gpurun_out/r2_scale2.log 2>&1
tail -80 gpurun_out/r2_scale2.log
