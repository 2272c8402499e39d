"""Time the diffusion-head sampler (one DiffLoss.sample = 100-step chain) against the eager port on the same GPU.

    python tools/diff_bench.py [rows ...]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
from vlg.diffloss import DiffLoss  # noqa: E402
from oracle import diffusion as D  # noqa: E402

dev = "cuda"
rows_list = [int(a) for a in sys.argv[1:]] or [1, 2, 8, 16]
torch.manual_seed(0)
d = DiffLoss(target_channels=8, z_channels=2048, depth=3, width=1024, num_sampling_steps="100").eval()
for m in d.modules():
    if isinstance(m, torch.nn.Linear):
        m.weight.data.normal_(0, 1.0 / m.weight.shape[1] ** 0.5)
        m.bias.data.normal_(0, 0.05)
d = d.to(dev)
w = {k: v for k, v in d.state_dict().items()}
for rows in rows_list:
    z = torch.randn(rows, 2048, device=dev)
    for _ in range(int(os.environ.get('DIFF_BENCH_WARM', '3'))):
        d.sample(z, 1.0, 1.0, seed=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get('DIFF_BENCH_ITERS', '20'))
    e0.record()
    for i in range(n):
        d.sample(z, 1.0, 1.0, seed=i)
    e1.record()
    torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / n
    if os.environ.get('DIFF_BENCH_NO_EAGER'):
        print(f'rows {rows:2d}: vlg {ours:.3f} ms per token', flush=True)
        continue
    noise = torch.randn(rows, 8, device=dev)
    sn = torch.randn(100, rows, 8, device=dev)
    D.p_sample_loop(w, z, noise, sn)
    torch.cuda.synchronize()
    t = time.time()
    D.p_sample_loop(w, z, noise, sn)
    torch.cuda.synchronize()
    eager = (time.time() - t) * 1e3
    print(f"rows {rows:2d}: vlg {ours:.3f} ms per token ({ours * 10:.1f} us per diffusion step), eager port {eager:.1f} ms "
          f"({eager / ours:.1f}x)", flush=True)
