#!/bin/bash
# full GPU test suite + bench (default path) + whole-trunk kernel A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-200
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; echo "bench exit $?"; cut -c1-1800 gpurun_out/bench_r11.json; tail -3 gpurun_out/bench_r11.err
VLG_DL=1 timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>&1 | tail -1 | cut -c1-250
