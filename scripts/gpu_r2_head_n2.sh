#!/bin/bash
# 2-GPU call: the headline problem (3-D Stokes IB, nel=74, 10.35 M DoFs) row-partitioned over 2 B200s, launched the
# way the driver launches the scaling run.  One warm-up and one timed solve: the point is the setup path (rank 0
# builds and cuts 1.8 G non-zeros, /dev/shm hand-over, per-rank finalize with the device BSR conversion) at the
# largest per-rank footprint, and the parity object of the partitioned solve.
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  free -g | head -2; df -h /dev/shm | tail -1
  FDAL_VERBOSE_SETUP=1 timeout 560 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29541 \
      bench.py --gpus 2 --steps 1 --warmup 1 2>gpurun_out/head_n2.err | tee gpurun_out/head_n2.json | cut -c1-400
  grep -E "bench |fdal_finalize\]|Error|error|Traceback|Killed" gpurun_out/head_n2.err | cut -c1-220 | tail -40
  free -g | head -2
} > gpurun_out/r2_head_n2.log 2>&1
tail -80 gpurun_out/r2_head_n2.log
