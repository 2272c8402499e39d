#!/bin/bash
# round-2 call B: full GPU suite after the right-sized activation box, VQ variants + ncu, small-row configs A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-250
for v in 0 1 2 3 4; do VLG_VQ_VARIANT=$v python tools/vq_bench.py 2>&1 | tail -1 | cut -c90-330 | sed "s/^/variant $v: /"; done
python tools/vq_bench.py > gpurun_out/vq_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vq_score -c 1 -f -o gpurun_out/r2_vq_score python tools/vq_bench.py > gpurun_out/vq_ncu.log 2>&1; tail -2 gpurun_out/vq_ncu.log
for fb in 1 0; do
  VLG_LT_FULLBOX=$fb python tools/cfg_sweep.py 2>&1 | grep cfg- | sed "s/^/FULLBOX=$fb /"
  VLG_LT_FULLBOX=$fb python tools/published_cfg.py 2>&1 | tail -5 | sed "s/^/FULLBOX=$fb /"
done
