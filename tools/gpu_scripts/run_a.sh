#!/bin/bash
# round-2 validation call A: full GPU suite, VQ bench (both paths), layer-tail switches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | cut -c1-220
python tools/vq_bench.py 2>&1 | tail -1
VLG_VQ_EXACT=1 python tools/vq_bench.py 2>&1 | tail -1
for s in 16 8; do for f in 0 1; do
  VLG_LT_MAXS=$s VLG_LT_FASTSILU=$f timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-140 | sed "s/^/MAXS=$s FASTSILU=$f /"
done; done
