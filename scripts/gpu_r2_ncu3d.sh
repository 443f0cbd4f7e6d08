#!/bin/bash
# 1-GPU call: ncu --set full of the fused Chebyshev SpMV and the plain SpMV on the 3-D workload (nel=40) + launch list
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  echo "== plain run first (must exit 0 without ncu)"
  timeout 900 python bench.py --workload stokes3d --nel 40 --steps 1 --warmup 1 --no-parity 2>gpurun_out/ncu3d_plain.err | cut -c1-300
  echo "== ncu --set full: k_bsr_spmv instances inside the kernel-timing loops (after the solves)"
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_bsr_spmv --launch-skip 100 --launch-count 10 \
     -o gpurun_out/ncu_full_bsr3_nel40 -f python bench.py --workload stokes3d --nel 40 --steps 1 --warmup 0 --no-parity > gpurun_out/ncu3d_full.log 2>&1
  tail -5 gpurun_out/ncu3d_full.log
  ncu -i gpurun_out/ncu_full_bsr3_nel40.ncu-rep --page raw --csv > gpurun_out/ncu_full_bsr3_nel40_raw.csv 2>/dev/null
  wc -l gpurun_out/ncu_full_bsr3_nel40_raw.csv
} > gpurun_out/r2_ncu3d.log 2>&1
tail -30 gpurun_out/r2_ncu3d.log
