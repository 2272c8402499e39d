#!/bin/bash
# Round-1 profiling evidence for the final kernels: launch list of one mid-sequence decode step + full captures.
mkdir -p gpurun_out
export VLG_PROFILE_STEP=286
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_step287.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc|attn_decode_tma|cfg_sample" -c 8 -o gpurun_out/r1_full_step287 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -1 gpurun_out/ncu_full.log
