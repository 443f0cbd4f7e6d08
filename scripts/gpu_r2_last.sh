#!/bin/bash
# last 1-GPU call of the round: smoke() and the two new suites on the final sources
cd /root/repo || exit 1
mkdir -p gpurun_out
{
  timeout 150 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
  timeout 150 python -m pytest tests/test_gpu_mass_cheb.py tests/test_gpu_bsr_build.py -q -k "not full_size" 2>&1 | grep -v Warning | tail -5
} > gpurun_out/r2_last.log 2>&1
tail -20 gpurun_out/r2_last.log
