set -x
timeout 70 python -m pytest tests/test_multi_gpu.py -x -q -k "stokes3d_diag or stokes2d_diag" > gpurun_out/pytest_2gpu.log 2>&1; tail -4 gpurun_out/pytest_2gpu.log
