#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4 | cut -c1-250
python tools/published_cfg.py 2>&1 | tail -4
python tools/cfg_sweep.py cfg-4 2>&1 | grep cfg-
timeout 900 python bench.py > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_i.json; tail -3 gpurun_out/bench_i.err
