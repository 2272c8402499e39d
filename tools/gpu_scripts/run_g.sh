#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decode.py -q -x 2>&1 | tail -3 | cut -c1-200
timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150
python tools/published_cfg.py 2>&1 | tail -4
python tools/cfg_sweep.py cfg-4 2>&1 | grep cfg-
python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt1_phases.txt
python tools/vq_bench.py > gpurun_out/vq_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vq_score -c 1 -f -o gpurun_out/r2_vq_score python tools/vq_bench.py > gpurun_out/vq_ncu.log 2>&1; tail -2 gpurun_out/vq_plain.log | cut -c1-300; tail -2 gpurun_out/vq_ncu.log | cut -c1-200
