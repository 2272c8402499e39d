#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attn" -x > gpurun_out/tests_attn2.log 2>&1; echo "attn tests exit $?"; tail -n 5 gpurun_out/tests_attn2.log
python tools/attn_sweep.py 2 > gpurun_out/attn_sweep_tma3.txt 2>&1; cat gpurun_out/attn_sweep_tma3.txt
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x > gpurun_out/tests_dec2.log 2>&1; echo "decode tests exit $?"; tail -n 3 gpurun_out/tests_dec2.log
python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_attn2.json 2> gpurun_out/bench_attn2.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_attn2.json; grep -o '"roofline".*' gpurun_out/bench_attn2.json | cut -c1-400
