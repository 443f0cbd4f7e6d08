set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for wl in stokes2d_diag stokes3d stokes2d_1M laplace; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 400 gpurun_out/bench_$wl.err
done
FDAL_NO_STREAM=1 timeout 900 python bench.py --workload stokes2d_diag --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_stokes2d_diag_nostream.json 2>&1
timeout 900 python bench.py --workload stokes2d_diag --steps 3 --warmup 3 --no-cpu --no-graphs > gpurun_out/bench_stokes2d_diag_nographs.json 2>&1
