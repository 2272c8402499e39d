#!/bin/bash
# ncu evidence for one change: launch list (shares) at mid-sequence + full capture of attention and GEMM kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 148610 -c 344 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:"attn_decode_tma|gemm_bf16_tc" -s 104544 -c 10 -o gpurun_out/prof_r1 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"; tail -2 gpurun_out/ncu_full.log; ls -la gpurun_out/
