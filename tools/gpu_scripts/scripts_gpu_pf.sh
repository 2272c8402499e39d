#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_decode.py -q -m gpu 2>&1 | tail -6 | cut -c1-250
timeout 200 python tools/prefill_ab.py 2>&1 | tail -2
VLG_STEPPED_PREFILL=1 timeout 200 python tools/prefill_ab.py 2>&1 | tail -2
