#!/bin/bash
# layer-tail kernel: parity tests, phase timings, then bench A/B against the separate-launch path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -15
VLG_LT_DEBUG=1 timeout 200 python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_phases.txt
timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras > gpurun_out/bench_lt.json 2> gpurun_out/bench_lt.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_lt.json; tail -3 gpurun_out/bench_lt.err
