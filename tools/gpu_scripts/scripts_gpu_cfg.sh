#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_decode.py -q -m gpu -k "batched_prefill" 2>&1 | tail -3 | cut -c1-200
timeout 900 python tools/cfg_sweep.py 2>&1 | grep "cfg-" | tee gpurun_out/cfg_sweep.txt
