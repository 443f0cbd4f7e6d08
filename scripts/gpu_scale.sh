set -x
mkdir -p gpurun_out
NG=${NG:-4}
nproc; free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/scale_default_n$NG.json 2> gpurun_out/scale_default_n$NG.err; tail -c 600 gpurun_out/scale_default_n$NG.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $NG --workload stokes3d --nel 30 --steps 2 --warmup 1 --no-cpu > gpurun_out/scale_s3d_n$NG.json 2> gpurun_out/scale_s3d_n$NG.err; tail -c 600 gpurun_out/scale_s3d_n$NG.err
timeout 600 python -m pytest tests/test_multi_gpu.py -q -x 2>&1 | tail -3
