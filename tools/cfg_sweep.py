"""The other BASELINE configs at full size on one GPU (they are parity-test cases, not bench lines): throughput and the
fraction of the per-config HBM roofline of SURVEY.md 8d.
  cfg-3  GPT-XXL c2i 384 px, the per-GPU shard of batch 256 on 8 GPUs (32 samples, 64 cache rows), 576 tokens
  cfg-4  GPT-XL t2v, 5 x 32 x 32 = 5120 latents, 120-token synthetic caption prefix with left-padded masks, CFG 7.5, B = 1 / 8
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg.gpt import GPT_models
from vlg.generate import generate
from vlg import gpt_video
from vlg.generate_video import generate as generate_video

ROOF = {"cfg-3": 25.4e3, "cfg-4 B=1": 2.66e3, "cfg-4 B=8": 5.59e3}     # tok/s at the measured 6 544 GB/s


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


which = sys.argv[1:] or ["cfg-3", "cfg-4"]
if "cfg-3" in which:
    torch.manual_seed(0)
    m = GPT_models["GPT-XXL"](block_size=576, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(m.output.weight, std=0.02)
    m = m.to("cuda", torch.bfloat16)
    cond = torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(1234)).cuda()
    dt, out = timed(lambda: generate(m, cond, 576, cfg_scale=4.0, temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True))
    assert out.shape == (32, 576) and int(out.min()) >= 0 and int(out.max()) < 16384
    v = 32 * 576 / dt
    print(f"cfg-3 GPT-XXL shard (32 samples, 64 rows): {dt * 1e3:8.1f} ms  {v:9.0f} tok/s per GPU  = {v / ROOF['cfg-3']:.3f} of the HBM roofline", flush=True)
    del m
    torch.cuda.empty_cache()
if "cfg-4" in which:
    torch.manual_seed(0)
    m = gpt_video.GPT_models["GPT-XL"](block_size=1024, cls_token_num=120, model_type="t2v", vae_embed_dim=8, num_frames=17,
                                       t_downsample_size=4, shuffle_video_tokens=False).eval()
    for n, p in m.named_parameters():
        if "vae_latent_adapter" in n:
            torch.nn.init.normal_(p, std=0.05)
    m = m.to("cuda", torch.bfloat16)
    g = torch.Generator().manual_seed(1234)
    for B in (1, 8):
        cond = torch.randn(B, 120, 2048, generator=g)
        masks = torch.zeros(B, 120)
        for b in range(B):
            masks[b, 120 - int(torch.randint(8, 121, (1,), generator=g)):] = 1
        cond = (cond * masks[:, :, None]).cuda().bfloat16()
        dt, out = timed(lambda: generate_video(m, cond, 5120, masks.cuda(), cfg_scale=7.5), reps=1)
        assert out.shape == (B, 5120, 8) and bool(torch.isfinite(out.float()).all())
        v = B * 5120 / dt
        print(f"cfg-4 GPT-XL t2v B={B} ({2 * B} rows, 5120 latents): {dt * 1e3:9.1f} ms  {v:8.0f} latents/s  = {v / ROOF[f'cfg-4 B={B}']:.3f} of the HBM roofline", flush=True)
