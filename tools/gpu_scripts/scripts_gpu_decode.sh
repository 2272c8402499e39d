#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "vq" > gpurun_out/ops_vq.log 2>&1; echo "== vq exit $?"; tail -n 3 gpurun_out/ops_vq.log
for k in c2i_tiny_fp32 python_driven t2i_tiny c2i_tiny_bf16 gptb gptl; do
  timeout 900 python -m pytest tests/test_gpu_decode.py -q -m gpu -k "$k" > "gpurun_out/dec_$k.log" 2>&1
  echo "== $k -> exit $?"; tail -n 6 "gpurun_out/dec_$k.log"
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 3 gpurun_out/smoke.log
