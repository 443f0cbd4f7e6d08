set -x
export PROBE_CONFIGS="FDAL_UNROLL=4;FDAL_UNROLL=4,FDAL_TPR=4;FDAL_UNROLL=4,FDAL_TPR=16;FDAL_UNROLL=1;FDAL_UNROLL=4,FDAL_TPR=2"
timeout 600 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -7
timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -7
timeout 600 python scripts/kernel_probe.py laplace 2>&1 | tail -7
export PROBE_CONFIGS="FDAL_UNROLL=4"
timeout 600 python scripts/kernel_probe.py stokes2d_diag > gpurun_out/probe_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 40 -c 4 -o gpurun_out/prof_spmv_r1 -f python scripts/kernel_probe.py stokes2d_diag > gpurun_out/ncu_probe.log 2>&1
tail -3 gpurun_out/ncu_probe.log
