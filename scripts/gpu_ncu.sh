set -x
export PROBE_CONFIGS="FDAL_BSR_UNROLL=1"
timeout 100 python scripts/kernel_probe.py stokes2d_diag > gpurun_out/probe_plain.log 2>&1 && \
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_bsr_spmv -s 30 -c 3 -o gpurun_out/prof_bsr_cheb_r1 -f python scripts/kernel_probe.py stokes2d_diag > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
