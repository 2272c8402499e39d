#!/bin/bash
mkdir -p gpurun_out
export VLG_PROFILE_STEP=286
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_bf16_tc|attn_decode_tma" -c 6 -o gpurun_out/prof_r1_v1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
