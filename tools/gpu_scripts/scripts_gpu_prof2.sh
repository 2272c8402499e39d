#!/bin/bash
# Profiling evidence for the layer-tail build: launch list of two mid-sequence decode steps + full captures (run from the repo root).
mkdir -p gpurun_out
export VLG_PROFILE_STEP=286
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_lt_launches_step287.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"layer_tail|attn_decode_tma" -c 6 -o gpurun_out/r1_lt_full_step287 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/ncu_full.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"cfg_sample|gemm_bf16_tc|gather_rows" -c 3 -o gpurun_out/r1_lt_full_tail_step287 -f $CMD > gpurun_out/ncu_full2.log 2>&1
echo "full2 exit $?"; tail -1 gpurun_out/ncu_full2.log
