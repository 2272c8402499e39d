#!/bin/bash
mkdir -p gpurun_out
( python tools/gemm_sweep.py
  VLG_GEMM_DEBUG=1 python tools/gemm_sweep.py
  VLG_GEMM_DEBUG=2 python tools/gemm_sweep.py
  VLG_GEMM_DEBUG=3 python tools/gemm_sweep.py
  VLG_GEMM_BN=16 python tools/gemm_sweep.py
  VLG_GEMM_BN=64 python tools/gemm_sweep.py
  VLG_GEMM_BN=128 python tools/gemm_sweep.py
  VLG_PDL=0 python tools/gemm_sweep.py ) > gpurun_out/gemm_sweep.txt 2>&1
cat gpurun_out/gemm_sweep.txt
