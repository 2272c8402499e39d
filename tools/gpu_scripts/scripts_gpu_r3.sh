#!/bin/bash
mkdir -p gpurun_out
export VLG_PROFILE_STEP=286
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv; tail -3 gpurun_out/plain.log
