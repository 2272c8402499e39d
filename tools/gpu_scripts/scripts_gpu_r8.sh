#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu -x > gpurun_out/tests_r8.log 2>&1; echo "tests exit $?"; tail -n 4 gpurun_out/tests_r8.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r8.json 2> gpurun_out/bench_r8.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_r8.json; grep -o '"roofline".*' gpurun_out/bench_r8.json | cut -c1-400
VLG_GEMM_DEBUG=4 python tools/gemm_phases.py
