set -x
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q -s 2>&1 | tail -30
