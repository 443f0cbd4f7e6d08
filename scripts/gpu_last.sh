set -x
timeout 150 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_last.log 2>&1; tail -3 gpurun_out/pytest_last.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
