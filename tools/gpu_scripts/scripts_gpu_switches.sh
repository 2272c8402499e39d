#!/bin/bash
# the documented A/B switches must keep passing the decode parity tests
for e in "VLG_LT=0" "VLG_ATTN=generic" "VLG_PDL=0" "VLG_STEPPED_PREFILL=1" "VLG_NO_GRAPH=1" "VLG_FUSED_NORM=1 VLG_LT=0"; do
  echo "== $e"; env $e timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -x 2>&1 | tail -1 | cut -c1-160
done
