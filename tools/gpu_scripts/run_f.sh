#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-250
VLG_LT_V=2 python tools/lt_phases.py 288 2>&1 | tail -30 | tee gpurun_out/lt2_phases.txt
