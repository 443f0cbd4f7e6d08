set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
for wl in stokes2d_diag stokes3d laplace elasticity; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 300 gpurun_out/bench_$wl.err
done
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 300 gpurun_out/bench_reference.err
FDAL_VERBOSE_SETUP=1 timeout 2400 python bench.py --workload stokes3d --nel 64 --steps 2 --warmup 1 > gpurun_out/bench_s3d_nel64.json 2> gpurun_out/bench_s3d_nel64.err; grep -E "setup\]|Error|error" gpurun_out/bench_s3d_nel64.err | tail -12; tail -c 400 gpurun_out/bench_s3d_nel64.json
