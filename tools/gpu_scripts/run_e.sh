#!/bin/bash
# layer_tail2 first run: GPU suite, then cfg-2 bench A/B against the lock-step kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | cut -c1-250
for v in 2 1; do
  VLG_LT_V=$v timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150 | sed "s/^/LT_V=$v /"
done
