set -x
timeout 600 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -12
timeout 600 python scripts/kernel_probe.py stokes3d 2>&1 | tail -12
timeout 600 python scripts/kernel_probe.py laplace 2>&1 | tail -12
