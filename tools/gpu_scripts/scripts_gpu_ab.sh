#!/bin/bash
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu -k "attn or gptl or tiny_bf16 or geometries" 2>&1 | tail -2 | cut -c1-200
echo "== old lib (commit 6912608)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib (q prefetch)"; $B 2>&1 | tail -1 | cut -c80-200
echo "== old lib (commit 6912608)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib (q prefetch)"; $B 2>&1 | tail -1 | cut -c80-200
timeout 120 python tools/attn_sweep.py 2 2>&1 | grep -E "pos +(0|63|127|255|319|575) "
