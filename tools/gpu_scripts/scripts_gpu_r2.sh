#!/bin/bash
mkdir -p gpurun_out
# 1. correctness with PDL launches
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu -x > gpurun_out/tests_pdl.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/tests_pdl.log
# 2. bench PDL on / off
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_pdl.json 2> gpurun_out/bench_pdl.err; echo "bench pdl exit $?"; cut -c1-260 gpurun_out/bench_pdl.json
VLG_PDL=0 python bench.py --steps 3 --warmup 2 --skip-extras > gpurun_out/bench_nopdl.json 2>&1; echo "bench nopdl exit $?"; cat gpurun_out/bench_nopdl.json
# 3. attention sweep
python tools/attn_sweep.py 2 > gpurun_out/attn_sweep_tma.txt 2>&1; cat gpurun_out/attn_sweep_tma.txt
python tools/attn_sweep.py 1 > gpurun_out/attn_sweep_generic.txt 2>&1; tail -4 gpurun_out/attn_sweep_generic.txt
# 4. launch list of one mid-sequence decode step (profiler range)
export VLG_PROFILE_STEP=286
CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"; tail -2 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches.csv
