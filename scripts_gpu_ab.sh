#!/bin/bash
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
echo "== old lib LT=0"; VLG_LT=0 VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib LT=0 qregs"; VLG_LT=0 VLG_ATTN_QSMEM=0 $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib LT=0 qsmem"; VLG_LT=0 $B 2>&1 | tail -1 | cut -c80-200
echo "== old lib"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_old.so $B 2>&1 | tail -1 | cut -c80-200
echo "== new lib qregs"; VLG_ATTN_QSMEM=0 $B 2>&1 | tail -1 | cut -c80-200
