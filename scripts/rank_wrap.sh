#!/bin/bash
# torchrun --no-python scripts/rank_wrap.sh bench.py ... : local rank 0 runs under ncu's single-pass duration
# metric (no replay, so the spinning exchange kernels cannot dead-lock), the other ranks run plain
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${NCU_SKIP:-12000} --launch-count ${NCU_COUNT:-400} \
       --csv --log-file gpurun_out/${NCU_OUT:-launches_rank0.csv} python "$@"
else
  exec python "$@"
fi
