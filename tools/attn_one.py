"""Run the TMA attention kernel a few times at one position (for ncu captures): python tools/attn_one.py POS"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 255
rows, H, D, S_max = 128, 16, 64, 584
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.randn(rows, H, D, device="cuda", generator=g).bfloat16()
k = torch.randn(rows, H, S_max, D, device="cuda", generator=g).bfloat16()
v = torch.randn(rows, H, S_max, D, device="cuda", generator=g).bfloat16()
out = torch.empty(rows, H * D, device="cuda", dtype=torch.bfloat16)
for _ in range(6):
    L.check(L.lib().vlg_attn_decode(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), rows, H, D, S_max, pos, None, 0, L.VLG_BF16, 2, None, 0, L.stream()))
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
