"""One fp32 trunk position of the GPT-1B t2v model at batch 1 between cudaProfilerStart/Stop (the other half of a
generate_video_diff token), for `ncu --profile-from-start off`:
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv python tools/trunk_prof_fp32.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
from vlg.gpt_video_diff import GPT_models
torch.manual_seed(0)
m = GPT_models["GPT-1B"](block_size=1024, cls_token_num=120, caption_dim=2048, vae_embed_dim=8, num_frames=1, model_type="t2v").eval().to("cuda")
m.setup_caches(1, 120 + 1024, None)
lat = torch.randn(1, 1, 8, device="cuda")
for i in range(3):
    m(video_latent=lat, cond_embed=None, input_pos=torch.tensor([600 + i]))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m(video_latent=lat, cond_embed=None, input_pos=torch.tensor([610]))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
