"""Phase timestamps of one tcgen05 GEMM CTA (needs VLG_GEMM_DEBUG=4)."""
import os, sys, ctypes as C
os.environ.setdefault("VLG_GEMM_DEBUG", "4")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L
lib = L.lib()
names = ["start", "setup done", "B prefetch issued", "pdl_wait done", "first full", "last mma issued", "acc ready", "epilogue done", "dealloc done"]
for name, R, N, K in [("qkv", 128, 3072, 1024), ("w2", 128, 1024, 2816), ("head", 128, 16384, 1024)]:
    x = torch.randn(R, K, device="cuda").bfloat16()
    ws = [torch.randn(N, K, device="cuda").bfloat16() * 0.05 for _ in range(4)]
    out = torch.empty(R, N, device="cuda", dtype=torch.bfloat16)
    for i in range(4):
        L.check(lib.vlg_linear(L.ptr(x), L.ptr(ws[i]), L.ptr(out), R, N, K, L.VLG_BF16, 0, L.stream()))
        torch.cuda.synchronize()
    ts = (C.c_ulonglong * 16)()
    L.check(lib.vlg_debug_gemm_timestamps(ts))
    t0 = ts[0]
    print(name, " ".join(f"{n}={ts[i]-t0}ns" for i, n in enumerate(names)))
