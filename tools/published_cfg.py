"""The workload behind the reference's only published throughput numbers (autoregressive/serve/README.md:10-16, A100):
8 class labels, CFG 4.0 (16 cache rows), 384 px = 576 tokens, top-k 2000, top-p 1.0, temperature 1.0, bf16 --
timed here on one B200 with random-init weights (BASELINE.md section 1 lists the A100 seconds)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg.gpt import GPT_models
from vlg.generate import generate

A100 = {"GPT-B": (7.80, 2.39), "GPT-L": (13.72, 3.48), "GPT-XL": (19.76, 4.84), "GPT-XXL": (26.38, 6.36)}   # eager s, vLLM s
cond = torch.randint(0, 1000, (8,), generator=torch.Generator().manual_seed(0)).cuda()
for name, (eager, vllm) in A100.items():
    torch.manual_seed(0)
    m = GPT_models[name](block_size=576, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(m.output.weight, std=0.02)
    m = m.to("cuda", torch.bfloat16)
    kw = dict(cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True)
    generate(m, cond, 576, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = generate(m, cond, 576, **kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name:8s} 8 images x 576 tokens: {dt:6.3f} s = {8 * 576 / dt:8.0f} tok/s on B200   (published A100: eager {eager:.2f} s, "
          f"vLLM {vllm:.2f} s -> x{eager / dt:.0f} / x{vllm / dt:.1f})", flush=True)
    del m
    torch.cuda.empty_cache()
