"""Microbenchmark: vlg_linear (tcgen05 GEMM) time for the decode-step shapes.  Env: VLG_GEMM_BN, VLG_GEMM_DEBUG."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L

dev = "cuda"
lib = L.lib()
st = torch.cuda.Stream()
shapes = [("qkv", 128, 3072, 1024), ("wo", 128, 1024, 1024), ("w13", 128, 5632, 1024), ("w2", 128, 1024, 2816),
          ("head", 128, 16384, 1024), ("qkv_r8", 8, 3072, 1024)]
tag = f"BN={os.environ.get('VLG_GEMM_BN','auto')} DBG={os.environ.get('VLG_GEMM_DEBUG','0')}"
with torch.cuda.stream(st):
    for name, R, N, K in shapes:
        x = torch.randn(R, K, device=dev).bfloat16()
        ws = [torch.randn(N, K, device=dev).bfloat16() * 0.05 for _ in range(24)]   # 24 weight sets: > L2 for every shape
        out = torch.empty(R, N, device=dev, dtype=torch.bfloat16)
        def call(i):
            L.check(lib.vlg_linear(L.ptr(x), L.ptr(ws[i % 24]), L.ptr(out), R, N, K, L.VLG_BF16, 0, L.stream()))
        call(0)
        st.synchronize()
        reps = 96
        gr = torch.cuda.CUDAGraph()          # graph replay: device time with PDL, free of the binding's host launch rate
        with torch.cuda.graph(gr, stream=st):
            for i in range(reps):
                call(i)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.synchronize()
        e0.record(st)
        gr.replay()
        e1.record(st)
        st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        mb = N * K * 2 / 1e6
        print(f"{tag:18s} {name:7s} R={R:3d} N={N:5d} K={K:4d}  {us:7.2f} us  weights {mb:5.1f} MB  {mb / us * 1e3:7.0f} GB/s", flush=True)
