set -x
FDAL_OVERLAP=1 timeout 200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_overlap.log 2>&1; tail -3 gpurun_out/pytest_overlap.log
FDAL_OVERLAP=1 timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_default_overlap.json 2> gpurun_out/bench_default_overlap.err; tail -c 200 gpurun_out/bench_default_overlap.err
