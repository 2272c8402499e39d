#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r10.json 2> gpurun_out/bench_r10.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_r10.json; tail -2 gpurun_out/bench_r10.err
VLG_FUSED_NORM=0 python bench.py --steps 3 --warmup 2 --skip-extras 2>&1 | tail -1
