#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_phases.py
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu -x > gpurun_out/tests_elect.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/tests_elect.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_elect.json 2> gpurun_out/bench_elect.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_elect.json; grep -o '"roofline".*' gpurun_out/bench_elect.json | cut -c1-400
python tools/attn_sweep.py 2 > gpurun_out/attn_sweep_tma2.txt 2>&1; cat gpurun_out/attn_sweep_tma2.txt
