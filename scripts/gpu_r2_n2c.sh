#!/bin/bash
cd /root/repo || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 $TR --nproc-per-node 2 --master-port $((29550 + RANDOM % 200)) bench.py --gpus 2 --workload stokes3d --nel 40 --steps 3 --warmup 2 --no-parity 2>gpurun_out/r2_n2c_$name.err > gpurun_out/r2_n2c_$name.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_n2c_$name.json").read().strip().splitlines()[-1]); s=d["solve"]
print("$name", "ms", round(d["ms_per_step"],1), "ms/inner", round(s["ms_per_inner_iteration"],4), {k:(round(v["ms"]*1e3,1),v["launches"]) for k,v in d["kernels"].items() if "ms" in v and k in ("spmv_A","aug_apply","vcycle")})
PY
}
{
  run p2p_default A=1
  run p2p_nosplit FDAL_NO_SPLIT=1
  run nccl FDAL_COMM=nccl
  run p2p_nographs FDAL_NO_DIST_GRAPHS=1
} > gpurun_out/r2_n2c.log 2>&1
cat gpurun_out/r2_n2c.log
