"""A/B of the batched T=120 prefill against stepping the prefix position by position (VLG_STEPPED_PREFILL=1):
t2i GPT-XL, batch 8 with CFG (16 rows), bf16; time of generate() for a short continuation, and token agreement."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg.gpt import GPT_models
from vlg.generate import generate
torch.manual_seed(0)
m = GPT_models["GPT-XL"](block_size=256, cls_token_num=120, model_type="t2i").eval()
torch.nn.init.normal_(m.output.weight, std=0.02)
m = m.to("cuda", torch.bfloat16)
g = torch.Generator().manual_seed(1)
cond = torch.randn(8, 120, 2048, generator=g).cuda().bfloat16()
masks = torch.zeros(8, 120)
for b in range(8):
    masks[b, 120 - (8 + 14 * b):] = 1
cond = cond * masks.cuda()[:, :, None].bfloat16()
for n_new in (1, 32):
    out = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = generate(m, cond, n_new, masks.cuda(), cfg_scale=7.5, temperature=1.0, top_k=1, top_p=1.0, sample_logits=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"stepped={os.environ.get('VLG_STEPPED_PREFILL', '0')} new_tokens={n_new}: {dt * 1e3:8.2f} ms  first tokens {out[:4, 0].tolist()}", flush=True)
