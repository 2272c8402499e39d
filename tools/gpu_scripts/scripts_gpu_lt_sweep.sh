#!/bin/bash
B="timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras"
echo "== default (384 threads, 6 stages)"; $B 2>&1 | tail -1 | cut -c80-200
for v in l320_6 l384_5 l384_4; do echo "== $v"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_$v.so $B 2>&1 | tail -1 | cut -c80-200; done
echo "== default"; $B 2>&1 | tail -1 | cut -c80-200
