#!/bin/bash
echo "== default lib"; timeout 300 python tools/two_stream.py 2>&1 | tail -2
echo "== small-footprint lib (attention 8 stages/8 slots, layer tail 256 thr/112 regs/4 stages)"; VLG_LIB_PATH=$PWD/video-llamagen_b200/lib/ab/libvlg_small.so timeout 300 python tools/two_stream.py 2>&1 | tail -2
