#!/bin/bash
# same-box A/B of the working tree against a reference build (video-llamagen_b200/lib/ab/libvlg_head.so = HEAD)
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_decode.py -q -m gpu 2>&1 | tail -2 | cut -c1-200
for i in 1 2; do
echo "== reference lib (HEAD)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_head.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib"; $B 2>&1 | tail -1 | cut -c80-200
done
