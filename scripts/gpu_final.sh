# final evidence run of the round (single GPU, bounded): tests, smoke, bench lines, ncu launch list + one full capture
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu > gpurun_out/pytest_final.log 2>&1; tail -4 gpurun_out/pytest_final.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 200 gpurun_out/bench_default.err
timeout 150 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 120 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/plain_bench.log 2>&1 && \
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_spmv|k_bsr|k_cg_|k_mass|k_dot|k_multi|k_axpby|k_gemv|k_cheb|k_diag|k_scale|k_couple|k_pack" -s 2000 -c 4000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_r1.csv
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_bsr_spmv.*EpiCheb" -s 60 -c 3 -o gpurun_out/prof_bsr_cheb_r1 -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
export PROBE_CONFIGS="FDAL_BSR_UNROLL=1;FDAL_BSR_UNROLL=2;FDAL_BSR_UNROLL=2,FDAL_BSR_TPR=2"
timeout 150 python scripts/kernel_probe.py stokes2d_diag 2>&1 | tail -4
