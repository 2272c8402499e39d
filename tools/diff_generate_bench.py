"""generate_video_diff.generate at the size the reference's script runs (GPT-1B trunk, fp32, batch 1, cfg_scale 1, 32 x 32
latent grid of one frame, 120-token caption prefix): ms per generated token, split into trunk position and diffusion head.

    python tools/diff_generate_bench.py [n_tokens]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
from vlg.gpt_video_diff import GPT_models  # noqa: E402
from vlg.generate_video_diff import generate  # noqa: E402

n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 48
torch.manual_seed(0)
m = GPT_models["GPT-1B"](block_size=1024, cls_token_num=120, caption_dim=2048, vae_embed_dim=8, num_frames=1,
                         model_type="t2v").eval().to("cuda")
cond = torch.randn(1, 120, 2048, device="cuda")
mask = torch.ones(1, 120, device="cuda")
generate(m, cond, 4 + n_tok, emb_masks=mask, temperature=1.0, cfg_iter=1.0)     # sizes the engine (KV cache, weight views) once
torch.cuda.synchronize()
t0 = time.perf_counter()
generate(m, cond, 4, emb_masks=mask, temperature=1.0, cfg_iter=1.0)
torch.cuda.synchronize()
t4 = time.perf_counter() - t0
t0 = time.perf_counter()
generate(m, cond, 4 + n_tok, emb_masks=mask, temperature=1.0, cfg_iter=1.0)
torch.cuda.synchronize()
tn = time.perf_counter() - t0
per_tok = (tn - t4) / n_tok * 1e3
# trunk position alone
eng = m._engine
lat = torch.randn(1, 1, 8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20):
    m(video_latent=lat, cond_embed=None, input_pos=torch.tensor([130 + i]))
torch.cuda.synchronize()
trunk = (time.perf_counter() - t0) / 20 * 1e3
z = torch.randn(1, 2048, device="cuda")
t0 = time.perf_counter()
for i in range(20):
    m.diffloss.sample(z, 1.0, 1.0)
torch.cuda.synchronize()
head = (time.perf_counter() - t0) / 20 * 1e3
w_gb = sum(p.numel() for n, p in m.named_parameters() if n.startswith("layers.")) * 4 / 1e9
print(f"GPT-1B fp32, batch 1: {per_tok:.2f} ms per token ({1e3 / per_tok:.0f} latents/s); prefill + 4 tokens {t4 * 1e3:.0f} ms; "
      f"trunk position {trunk:.2f} ms ({w_gb:.2f} GB of weights: {w_gb / trunk * 1e3:.0f} GB/s), diffusion head {head:.2f} ms")
