set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in stokes2d_1M stokes2d_diag stokes3d; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 300 gpurun_out/bench_$wl.err
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu --no-bsr > gpurun_out/bench_${wl}_nobsr.json 2> gpurun_out/bench_${wl}_nobsr.err; tail -c 300 gpurun_out/bench_${wl}_nobsr.err
done
FDAL_VERBOSE_SETUP=1 timeout 1500 python bench.py --workload stokes3d --nel 48 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_s3d_nel48.json 2> gpurun_out/bench_s3d_nel48.err; grep -E "setup\]|Error|error" gpurun_out/bench_s3d_nel48.err | tail -25; tail -c 300 gpurun_out/bench_s3d_nel48.json
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
