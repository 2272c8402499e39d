"""cfg-5: VQ argmin over 1M latents (1024x8x32x32) against the 16384x8 l2-normalised codebook: latents/s, FP32 TFLOP/s and
the fraction of the FP32 FMA peak measured on this GPU (register-resident FMA loop, scalar and packed)."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
import torch
from vlg import _lib as L
from vlg.vq_model import VectorQuantizer


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fp32_peak():
    out = torch.zeros(4, device="cuda")
    res = {}
    for variant, name in ((0, "ffma"), (1, "ffma2")):
        fl = C.c_longlong()
        ms = timed(lambda: L.check(L.lib().vlg_debug_fp32_fma_peak(variant, 4096, L.ptr(out), C.byref(fl), L.stream())), 5)
        res[name] = round(fl.value / ms / 1e9, 2)
    return res


if __name__ == "__main__":
    torch.manual_seed(0)
    vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval().cuda()
    z = torch.randn(1024, 8, 32, 32, device="cuda")
    ms = timed(lambda: vq(z), 10)
    n = z.shape[0] * z.shape[2] * z.shape[3]
    peak = fp32_peak()
    tf = 2 * n * 16384 * 8 / ms / 1e9
    print(json.dumps({"workload": "cfg-5 vq argmin 1,048,576 latents x 16384 codes x 8 dims (fp32, index-exact)", "ms": round(ms, 3),
                      "latents_per_s": round(n / ms * 1e3), "fp32_tflops": round(tf, 2), "fp32_peak_tflops_measured": peak,
                      "frac_of_fp32_peak": round(tf / max(peak.values()), 4), "path": os.environ.get("VLG_VQ_EXACT", "0") == "1" and "exact kernel" or "score kernel + exact recheck",
                      "note": "includes codebook normalisation prep kernel, z_q gather and the recheck pass"}))
