#!/bin/bash
mkdir -p gpurun_out
VLG_LT_DEBUG=1 timeout 200 python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_phases.txt
