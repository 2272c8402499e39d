#!/bin/bash
# layer-tail kernel: parity tests (incl. the geometry sweep), per-phase timings of the middle of a cfg-2 decode
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -q -m gpu 2>&1 | tail -12 | cut -c1-220
VLG_LT_DEBUG=1 timeout 200 python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_phases.txt
