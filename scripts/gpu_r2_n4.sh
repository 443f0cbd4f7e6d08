#!/bin/bash
# 4-GPU call: the multi-GPU parity suite (2- and 4-rank cases, both communication modes), then 3-D strong scaling at N=4
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  echo "== multi-GPU parity"
  timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | grep -v Warning | tail -8
  echo "== 3-D nel=48, N=4 strong"
  timeout 900 $TR --nproc-per-node 4 --master-port 29531 bench.py --gpus 4 --workload stokes3d --nel 48 --steps 3 --warmup 2 2>gpurun_out/r2_s3d48_n4.err | tee gpurun_out/r2_s3d48_n4.json | cut -c1-300
  grep -E "bench |Error|error|Traceback" gpurun_out/r2_s3d48_n4.err | tail -14
  echo "== 3-D nel=48, N=1 (same problem)"
  timeout 900 python bench.py --workload stokes3d --nel 48 --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_s3d48_n1.err | tee gpurun_out/r2_s3d48_n1.json | cut -c1-300
} > gpurun_out/r2_n4.log 2>&1
tail -60 gpurun_out/r2_n4.log
