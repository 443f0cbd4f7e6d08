#!/bin/bash
# last call of the round: bench lines of the two remaining BASELINE workloads on the final sources
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  timeout 110 python bench.py --workload elasticity --steps 3 --warmup 3 2>gpurun_out/last_elasticity.err | tee gpurun_out/last_elasticity.json | cut -c1-200
  timeout 70 python bench.py --workload laplace --steps 3 --warmup 3 2>gpurun_out/last_laplace.err | tee gpurun_out/last_laplace.json | cut -c1-200
  grep -h "bench " gpurun_out/last_*.err | cut -c1-160
} > gpurun_out/r2_last2.log 2>&1
tail -20 gpurun_out/r2_last2.log
