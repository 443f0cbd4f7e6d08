#!/bin/bash
# 8-GPU call: the 3-D Stokes IB problem (nel=48, 2.86 M DoFs) row-partitioned over 8 B200s (strong scaling vs the N=1/N=4 runs)
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  echo "== 3-D nel=48, N=8 strong"
  FDAL_BENCH_WATCHDOG_S=500 timeout 600 $TR --nproc-per-node 8 --master-port 29561 bench.py --gpus 8 --workload stokes3d --nel 48 --steps 3 --warmup 2 2>gpurun_out/r2_s3d48_n8.err | tee gpurun_out/r2_s3d48_n8.json | cut -c1-300
  grep -E "bench |Error|error|Traceback" gpurun_out/r2_s3d48_n8.err | tail -20
} > gpurun_out/r2_n8.log 2>&1
tail -40 gpurun_out/r2_n8.log
