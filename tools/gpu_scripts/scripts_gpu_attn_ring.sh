#!/bin/bash
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
echo "== default (16 stages, 16 slots)"; $B 2>&1 | tail -1 | cut -c80-200
for v in a24_8 a16_8 a8_16; do echo "== $v"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_$v.so $B 2>&1 | tail -1 | cut -c80-200; done
echo "== default"; $B 2>&1 | tail -1 | cut -c80-200
