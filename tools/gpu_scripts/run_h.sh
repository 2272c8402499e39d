#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "vq or VQ or codebook" 2>&1 | tail -3 | cut -c1-200
for v in 5 6 7 8; do VLG_VQ_VARIANT=$v python tools/vq_bench.py 2>&1 | tail -1 | cut -c90-290 | sed "s/^/variant $v: /"; done
