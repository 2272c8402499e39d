#!/bin/bash
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_ops.py -q -m gpu -k "decode or attn or bf16 or gptl" 2>&1 | tail -2 | cut -c1-200
echo "== old lib (commit 6912608)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib (q in smem)"; $B 2>&1 | tail -1 | cut -c80-200
echo "== old lib (commit 6912608)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib (q in smem)"; $B 2>&1 | tail -1 | cut -c80-200
