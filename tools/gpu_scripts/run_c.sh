#!/bin/bash
# round-2 call C: VQ variants, attention split model on the small-row configs, full bench line with the extras
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | cut -c1-250
for v in 0 5 2 3; do VLG_VQ_VARIANT=$v python tools/vq_bench.py 2>&1 | tail -1 | cut -c90-290 | sed "s/^/variant $v: /"; done
python tools/published_cfg.py 2>&1 | tail -4
python tools/cfg_sweep.py cfg-4 2>&1 | grep cfg-
timeout 900 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench exit $?"; cut -c1-6000 gpurun_out/bench_c.json; tail -3 gpurun_out/bench_c.err
