#!/bin/bash
# 1-GPU call: full -m gpu suite (new: Chebyshev mass solves, device AL-term scatter, 32-register BSR-3 kernels),
# then the bench lines the changes are about: elliptic (configs[2]) with the three forms of the exact M^-1,
# configs[1] (exact Mp^-1 in Chebyshev form), 3-D nel=40 (launch bound), and the ncu launch list of the 3-D run.
cd /root/repo || exit 1
mkdir -p gpurun_out
export FDAL_PARITY_TAG=1gpu_b
rm -f gpurun_out/parity_log_1gpu_b.jsonl
{
  echo "== new gpu tests first"
  timeout 900 python -m pytest tests/test_gpu_mass_cheb.py tests/test_gpu_setup.py -q --maxfail=6 2>&1 | grep -v Warning | tail -25
  echo "== elliptic cycle 6: default (persistent Chebyshev) / one kernel per iteration / Jacobi-PCG"
  timeout 400 python bench.py --workload elliptic --cycle 6 --steps 2 --warmup 1 2>gpurun_out/cb_elliptic_default.err | tee gpurun_out/cb_elliptic_default.json | cut -c1-160
  FDAL_MASS_CHEB=1 timeout 400 python bench.py --workload elliptic --cycle 6 --steps 2 --warmup 1 --no-parity 2>gpurun_out/cb_elliptic_cheb1.err | tee gpurun_out/cb_elliptic_cheb1.json | cut -c1-160
  FDAL_MASS_CHEB=0 timeout 400 python bench.py --workload elliptic --cycle 6 --steps 2 --warmup 1 --no-parity 2>gpurun_out/cb_elliptic_pcg.err | tee gpurun_out/cb_elliptic_pcg.json | cut -c1-160
  echo "== elliptic cycle 5 (m = 5 249): single-CTA PCG (default) vs persistent Chebyshev"
  timeout 300 python bench.py --workload elliptic --cycle 5 --steps 2 --warmup 1 --no-parity 2>gpurun_out/cb_elliptic5_default.err | tee gpurun_out/cb_elliptic5_default.json | cut -c1-160
  FDAL_MASS_CHEB_MIN_ROWS=1000 timeout 300 python bench.py --workload elliptic --cycle 5 --steps 2 --warmup 1 --no-parity 2>gpurun_out/cb_elliptic5_cheb.err | tee gpurun_out/cb_elliptic5_cheb.json | cut -c1-160
  echo "== configs[1]"
  timeout 400 python bench.py --workload stokes2d_1M --steps 3 --warmup 2 2>gpurun_out/cb_s2d1M.err | tee gpurun_out/cb_s2d1M.json | cut -c1-160
  echo "== 3-D nel=40"
  timeout 600 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-parity 2>gpurun_out/cb_s3d40.err | tee gpurun_out/cb_s3d40.json | cut -c1-160
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 --launch-count 400 --csv \
      --log-file gpurun_out/cb_launches_s3d40.csv python bench.py --workload stokes3d --nel 40 --steps 1 --warmup 0 --no-parity > gpurun_out/cb_ncu_list.log 2>&1
  wc -l gpurun_out/cb_launches_s3d40.csv
  echo "== whole gpu suite"
  timeout 1200 python -m pytest tests -m gpu -q --maxfail=8 --deselect tests/test_gpu_mass_cheb.py --deselect tests/test_gpu_setup.py 2>&1 | grep -v Warning | tail -15
  grep -h "bench " gpurun_out/cb_*.err | cut -c1-200
} > gpurun_out/r2_cheb.log 2>&1
tail -120 gpurun_out/r2_cheb.log
