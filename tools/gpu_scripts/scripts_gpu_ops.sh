#!/bin/bash
# per-kernel-family GPU test runs, each in its own process so a faulting kernel does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for k in vq rmsnorm linear_bf16 linear_fp32 attn_decode_bf16 attn_decode_fp32 "sample or top_k or cfg_mix or philox"; do
  name=$(echo "$k" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "$k" > "gpurun_out/ops_$name.log" 2>&1
  echo "== $k -> exit $?"; tail -n 4 "gpurun_out/ops_$name.log"
done
