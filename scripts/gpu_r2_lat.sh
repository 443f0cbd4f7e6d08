#!/bin/bash
# 2-GPU call: where does the per-exchange latency go?  (launch list of rank 0 under ncu's duration metric)
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
  echo "== N=1 kernel probe after the single-load-path fix"
  PROBE_CONFIGS="FDAL_BSR_TPR=4;FDAL_BSR_TPR=4,FDAL_BSR_UNROLL=1" timeout 600 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -3
  PROBE_CONFIGS="FDAL_BSR_TPR=16" timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -2
  echo "== 2 ranks, rank 0 under ncu (durations only), peer channels, no graphs"
  NCU_SKIP=14000 NCU_COUNT=500 NCU_OUT=lat_p2p_rank0.csv timeout 900 $TR --nproc-per-node 2 --master-port 29521 --no-python scripts/rank_wrap.sh bench.py --gpus 2 --workload stokes2d_diag --scaling weak --steps 1 --warmup 1 --no-parity --no-graphs 2>gpurun_out/lat_p2p.err | cut -c1-200
  tail -5 gpurun_out/lat_p2p.err
  echo "== same, NCCL fallback"
  FDAL_COMM=nccl NCU_SKIP=14000 NCU_COUNT=500 NCU_OUT=lat_nccl_rank0.csv timeout 900 $TR --nproc-per-node 2 --master-port 29522 --no-python scripts/rank_wrap.sh bench.py --gpus 2 --workload stokes2d_diag --scaling weak --steps 1 --warmup 1 --no-parity --no-graphs 2>gpurun_out/lat_nccl.err | cut -c1-200
  echo "== plain 2-rank runs: peer channels with FDAL_REP_ROWS=200000, graphs on"
  FDAL_REP_ROWS=200000 timeout 900 $TR --nproc-per-node 2 --master-port 29523 bench.py --gpus 2 --workload stokes2d_diag --scaling weak --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_s2d_n2_rep.err | tee gpurun_out/r2_s2d_n2_rep.json | cut -c1-200
} > gpurun_out/r2_lat.log 2>&1
tail -40 gpurun_out/r2_lat.log
