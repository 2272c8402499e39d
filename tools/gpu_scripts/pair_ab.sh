timeout 600 python -m pytest tests/test_gpu_decode.py -q -x 2>&1 | tail -4 | cut -c1-250
timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150
VLG_LT_PAIR=0 timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150 | sed 's/^/PAIR=0 /'
timeout 200 python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_pair_phases.txt
