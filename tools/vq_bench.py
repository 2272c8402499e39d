"""cfg-5: VQ argmin over 1M latents (1024x8x32x32) against the 16384x8 l2-normalised codebook: latents/s and FP32 TFLOP/s."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg.vq_model import VectorQuantizer
torch.manual_seed(0)
vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval().cuda()
z = torch.randn(1024, 8, 32, 32, device="cuda")
for _ in range(3):
    vq(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 10
for _ in range(reps):
    _, _, (_, _, idx) = vq(z)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n = z.shape[0] * z.shape[2] * z.shape[3]
print(json.dumps({"workload": "cfg-5 vq argmin 1,048,576 latents x 16384 codes x 8 dims (fp32, exact)", "ms": round(ms, 3),
                  "latents_per_s": round(n / ms * 1e3), "fp32_tflops": round(2 * n * 16384 * 8 / ms / 1e9, 2),
                  "note": "includes codebook normalisation prep kernel and z_q gather"}))
