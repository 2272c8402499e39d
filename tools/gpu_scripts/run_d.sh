#!/bin/bash
# round-2 call D (re-entry): full GPU suite, full bench line, launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench exit $?"; cut -c1-8000 gpurun_out/bench_d.json; tail -3 gpurun_out/bench_d.err
