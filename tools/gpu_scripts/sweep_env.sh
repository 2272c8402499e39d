#!/bin/bash
# A/B one environment switch on the profiling mode of the bench (device-resident generate() only):
#   tools/gpu_scripts/sweep_env.sh VLG_KV_PREFETCH_MB 0 48 80 112
# prints ms per 576-token generate for every value; run from the repo root under gpurun.
var=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras > gpurun_out/sweep_${var}_$v.json 2> gpurun_out/sweep_${var}_$v.err
  echo "$var=$v exit $? $(cut -c1-200 gpurun_out/sweep_${var}_$v.json)"
done
