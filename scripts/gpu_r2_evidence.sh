#!/bin/bash
# 1-GPU call: evidence for profiles/ — full -m gpu suite (parity log), ncu launch list + --set full capture on the 3-D
# workload, bench lines of the other BASELINE workloads (elliptic beta_2 sweep, laplace as shipped, configs[1], elasticity)
cd /root/repo || exit 1
mkdir -p gpurun_out
export FDAL_PARITY_TAG=1gpu
rm -f gpurun_out/parity_log_1gpu.jsonl
{
  echo "== gpu tests"
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -6
  echo "== plain 3-D nel=40 run (must exit 0 before ncu)"
  timeout 900 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-parity 2>gpurun_out/ev_s3d40.err | tee gpurun_out/ev_s3d40.json | cut -c1-200
  echo "== ncu launch list (durations) of the same command"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 --launch-count 400 --csv \
      --log-file gpurun_out/ev_launches_s3d40.csv python bench.py --workload stokes3d --nel 40 --steps 1 --warmup 0 --no-parity > gpurun_out/ev_ncu_list.log 2>&1
  tail -2 gpurun_out/ev_ncu_list.log | cut -c1-200; wc -l gpurun_out/ev_launches_s3d40.csv
  echo "== ncu --set full: 10 consecutive k_bsr_spmv launches inside the solve"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_bsr_spmv --launch-skip 100 --launch-count 10 \
      -o gpurun_out/ev_ncu_full_bsr3_nel40 -f python bench.py --workload stokes3d --nel 40 --steps 1 --warmup 0 --no-parity > gpurun_out/ev_ncu_full.log 2>&1
  tail -3 gpurun_out/ev_ncu_full.log | cut -c1-200
  ncu -i gpurun_out/ev_ncu_full_bsr3_nel40.ncu-rep --page raw --csv > gpurun_out/ev_ncu_full_bsr3_nel40_raw.csv 2>/dev/null; wc -l gpurun_out/ev_ncu_full_bsr3_nel40_raw.csv
  echo "== other BASELINE workloads"
  for b2 in 10 100 1000 10000 1000000; do
    timeout 600 python bench.py --workload elliptic --cycle 6 --beta2 $b2 --steps 3 --warmup 1 --no-parity 2>gpurun_out/ev_elliptic_$b2.err | tee gpurun_out/ev_elliptic_b$b2.json | cut -c1-120
  done
  timeout 600 python bench.py --workload laplace --steps 3 --warmup 2 2>gpurun_out/ev_laplace.err | tee gpurun_out/ev_laplace.json | cut -c1-120
  timeout 600 python bench.py --workload stokes2d_1M --steps 3 --warmup 2 2>gpurun_out/ev_s2d1M.err | tee gpurun_out/ev_s2d1M.json | cut -c1-120
  timeout 900 python bench.py --workload elasticity --steps 3 --warmup 1 --no-parity 2>gpurun_out/ev_elasticity.err | tee gpurun_out/ev_elasticity.json | cut -c1-120
} > gpurun_out/r2_evidence.log 2>&1
tail -60 gpurun_out/r2_evidence.log
