"""Per-phase timing of the layer-tail kernel inside a real cfg-2 decode, over all CTAs (VLG_LT_DEBUG=1)."""
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
torch.manual_seed(0)
m = GPT_models[os.environ.get("VLG_PH_MODEL", "GPT-L")](block_size=576, cls_token_num=1, model_type="c2i").eval()
torch.nn.init.normal_(m.output.weight, std=0.02)
m = m.to("cuda", torch.bfloat16)
cond = torch.randint(0, 1000, (int(os.environ.get("VLG_PH_BATCH", "64")),)).cuda()
n_new = int(sys.argv[1]) if len(sys.argv) > 1 else 288
names = {0: "start", 1: "pdl_wait", 2: "wo issued", 3: "bar", 4: "norm1", 5: "bar", 6: "w13 issued", 7: "bar", 8: "w2 issued",
         9: "bar", 10: "norm2", 11: "bar", 12: "qkv issued", 16: "wo acc", 17: "wo epi", 18: "w13 acc", 19: "w13 epi",
         20: "w2 acc", 21: "w2 epi", 22: "qkv acc", 23: "qkv epi",
         24: "w13 epi: acc 0 read", 25: "w13 epi: acc 1 added", 26: "w13 epi: math done"}
for rep in range(2):
    generate(m, cond, n_new, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, seed=rep)
    torch.cuda.synchronize()
NS = 32
ts = (C.c_ulonglong * (160 * NS))()
L.check(lib.vlg_debug_lt_timestamps(ts))
t = np.array(ts, dtype=np.int64).reshape(160, NS)[:148].astype(np.float64)
t0 = t[:, 1][t[:, 1] > 0].min()          # first CTA past pdl_wait
print(f"new_tokens {n_new}; times in us relative to the first CTA leaving pdl_wait; min / median / max over CTAs that have the stamp")
order = sorted(names)
for k in order:
    col = t[:, k]
    col = col[col > 0] - t0
    if len(col):
        print(f"  {names[k]:>22s}: {col.min()/1e3:7.2f} {np.median(col)/1e3:7.2f} {col.max()/1e3:7.2f}   (n={len(col)})")
