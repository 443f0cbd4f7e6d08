#!/bin/bash
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  echo "== 3-D nel=40, N=2 strong (boundary-first proportional CTA split)"
  timeout 900 $TR --nproc-per-node 2 --master-port 29541 bench.py --gpus 2 --workload stokes3d --nel 40 --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_s3d40_n2b.err | tee gpurun_out/r2_s3d40_n2b.json | cut -c1-300
  grep -E "bench |Error|error|Traceback" gpurun_out/r2_s3d40_n2b.err | tail -6
  echo "== multi-GPU parity (2-rank p2p cases)"
  timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "2-p2p" 2>&1 | grep -v Warning | tail -4
} > gpurun_out/r2_n2b.log 2>&1
tail -30 gpurun_out/r2_n2b.log
