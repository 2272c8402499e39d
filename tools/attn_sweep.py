"""Microbenchmark: attention kernel time vs cache position (fixed overhead and marginal bandwidth).
   python tools/attn_sweep.py [impl] [regions]   (impl 2 = TMA kernel, 1 = generic)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L

impl = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows, H, D, S_max = 128, 16, 64, 584
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
regions = int(sys.argv[2]) if len(sys.argv) > 2 else 6     # 1 = the same K/V region every launch (L2-hot when it fits)
q = torch.randn(rows, H, D, device=dev, generator=g).bfloat16()
ks = [torch.randn(rows, H, S_max, D, device=dev, generator=g).bfloat16() for _ in range(regions)]
vs = [torch.randn(rows, H, S_max, D, device=dev, generator=g).bfloat16() for _ in range(regions)]
out = torch.empty(rows, H * D, device=dev, dtype=torch.bfloat16)
nb = L.lib().vlg_attn_scratch_bytes(rows, H, D, S_max)
scratch = torch.empty(nb, dtype=torch.uint8, device=dev)
st = torch.cuda.Stream()
lib = L.lib()
with torch.cuda.stream(st):
    def call(pos, r):
        L.check(lib.vlg_attn_decode(L.ptr(q), L.ptr(ks[r]), L.ptr(vs[r]), L.ptr(out), rows, H, D, S_max, pos, None, 0,
                                    L.VLG_BF16, impl, L.ptr(scratch), nb, L.stream()))
    call(0, 0)
    st.synchronize()
    for pos in (0, 31, 63, 127, 191, 255, 319, 383, 447, 511, 575):
        # replay a CUDA graph of 60 launches: the host launch rate (~11 us per ctypes call) must not pollute the timing
        reps = 60
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for i in range(reps):
                call(pos, i % regions)
        gr.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.synchronize()
        e0.record(st)
        gr.replay()
        e1.record(st)
        st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        mb = rows * H * (pos + 1) * D * 2 * 2 / 1e6
        print(f"impl {impl} pos {pos:4d}  {us:7.2f} us  {mb:7.1f} MB  {mb / us * 1e3:7.1f} GB/s", flush=True)
