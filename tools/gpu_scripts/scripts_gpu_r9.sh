#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attn" -x 2>&1 | tail -2
python tools/attn_sweep.py 2 2>&1 | tee gpurun_out/attn_sweep_tma4.txt
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -2
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r9.json 2> gpurun_out/bench_r9.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_r9.json; grep -o '"roofline".*' gpurun_out/bench_r9.json | cut -c1-300
