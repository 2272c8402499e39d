#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
VLG_ATTN=generic python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_generic_attn.json 2>&1; echo "generic exit $?"; cat gpurun_out/bench_generic_attn.json
