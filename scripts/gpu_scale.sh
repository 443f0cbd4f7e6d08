set -x
mkdir -p gpurun_out
NG=${NG:-2}
timeout 900 python bench.py --workload stokes2d_diag --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_s2diag_n1.json 2> gpurun_out/scale_n1.err; tail -c 300 gpurun_out/scale_n1.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --workload stokes2d_diag --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_s2diag_n$NG.json 2> gpurun_out/scale_n$NG.err; tail -c 1500 gpurun_out/scale_n$NG.err
FDAL_DIST_GRAPHS=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --workload stokes2d_diag --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_s2diag_n${NG}_graphs.json 2> gpurun_out/scale_n${NG}_graphs.err; tail -c 800 gpurun_out/scale_n${NG}_graphs.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_default_n$NG.json 2> gpurun_out/scale_default_n$NG.err; tail -c 800 gpurun_out/scale_default_n$NG.err
timeout 900 python bench.py --workload stokes3d --nel 40 --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_s3d_nel40.json 2> gpurun_out/bench_s3d_nel40.err; tail -c 300 gpurun_out/bench_s3d_nel40.err
