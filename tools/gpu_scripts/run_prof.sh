#!/bin/bash
# ncu evidence for the round: launch list of two mid-sequence decode steps, full captures of the two step kernels and of the
# VQ score kernel.  Every ncu command follows a plain run of the same command line that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
export VLG_PROFILE_STEP=286
$CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_step287.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/prof_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"layer_tail_kernel|attn_decode_tma" -c 6 -f -o gpurun_out/r2_ncu_full_step287 $CMD > gpurun_out/prof_ncu2.log 2>&1
echo "full capture rc=$?"
unset VLG_PROFILE_STEP
python tools/vq_bench.py > gpurun_out/vq_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vq_score -c 1 -f -o gpurun_out/r2_vq_score python tools/vq_bench.py > gpurun_out/vq_ncu.log 2>&1
echo "vq capture rc=$?"; tail -1 gpurun_out/vq_plain.log | cut -c1-300
ls -la gpurun_out | tail -12
