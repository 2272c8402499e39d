"""Per-phase timing of the 1-2 row GEMV layer tail inside a real decode (VLG_LT_DEBUG=1): one sample with guidance = 2 rows.
    python tools/gv_phases.py [model] [new_tokens]"""
import os, sys, ctypes as C
os.environ.setdefault("VLG_LT_DEBUG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import numpy as np
import torch
from vlg import _lib as L
from vlg.gpt import GPT_models
from vlg.generate import generate
lib = L.lib()
name = sys.argv[1] if len(sys.argv) > 1 else "GPT-XL"
n_new = int(sys.argv[2]) if len(sys.argv) > 2 else 288
torch.manual_seed(0)
m = GPT_models[name](block_size=576, cls_token_num=1, model_type="c2i").eval()
torch.nn.init.normal_(m.output.weight, std=0.02)
m = m.to("cuda", torch.bfloat16)
cond = torch.randint(0, 1000, (1,)).cuda()
names = ["start", "pdl_wait", "wo done", "bar", "norm1 done", "w13 done", "bar", "acts staged", "w2 done", "bar", "norm2 done", "qkv done"]
for rep in range(2):
    generate(m, cond, n_new, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, seed=rep)
    torch.cuda.synchronize()
ts = (C.c_ulonglong * (160 * 32))()
L.check(lib.vlg_debug_gv_timestamps(ts))
t = np.array(ts, dtype=np.int64).reshape(160, 32)[:148].astype(np.float64)
t0 = t[:, 1][t[:, 1] > 0].min()
print(f"{name}, 2 rows, new_tokens {n_new}; us relative to the first CTA leaving pdl_wait; min / median / max over CTAs (warp 0)")
for k, nm in enumerate(names):
    col = t[:, k]
    col = col[col > 0] - t0
    if len(col):
        print(f"  {nm:>12s}: {col.min()/1e3:7.2f} {np.median(col)/1e3:7.2f} {col.max()/1e3:7.2f}   (n={len(col)})")
