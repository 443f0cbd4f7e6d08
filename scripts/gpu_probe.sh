set -x
export PROBE_CONFIGS="FDAL_NO_BSR=1;FDAL_BSR_TPR=8;FDAL_BSR_TPR=16;FDAL_BSR_TPR=4;FDAL_BSR_AOS=1,FDAL_BSR_TPR=8;FDAL_BSR_AOS=1,FDAL_BSR_TPR=16;FDAL_BSR_AOS=1,FDAL_BSR_TPR=4"
timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -9
timeout 600 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -9
