#!/bin/bash
mkdir -p gpurun_out
VLG_PH_BATCH=8 python tools/lt_phases.py 288 2>&1 | tail -22 | tee gpurun_out/lt_phases_b8.txt
VLG_PH_MODEL=GPT-XXL VLG_PH_BATCH=32 python tools/lt_phases.py 288 2>&1 | tail -22 | tee gpurun_out/lt_phases_xxl_b32.txt
