#!/usr/bin/env python
"""bench.py -- generated visual tokens/sec of the decode hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (config.workload): BASELINE cfg-2 -- GPT-L c2i 384 px (24x24 = 576 tokens), bf16, batch 64 with CFG 4.0
(128 cache rows), top-k 2000, temperature 1.0, random-init weights, synthetic class labels.  One "step" is one
full generate() over one batch: prefill + 575 graph-replayed decode steps.

  value       tokens/s with the class labels already resident in HBM (vlg_generate), CUDA events, max over ranks
  e2e         same through the host-buffer C-ABI call (vlg_generate_c2i_host): pinned labels H2D + tokens D2H inside
  roofline    the kernel with the largest measured share of the step (layer_tail_kernel or attn_decode_tma_kernel, both
              timed alone with CUDA events on their stream); the other one and the whole-step figure ride along
  cpu_baseline  the oracle port (reference algorithm restated, oracle/ar_decode.py) on this host's cores, bounded sample
  extras      (N = 1 only) the other BASELINE configs at full size -- cfg3 (GPT-XXL shard), cfg4_b8 (GPT-XL t2v), cfg5_vq
              (1 M-latent index search against the measured FP32 FMA peak) -- decode_code, diff_head (the diffusion-head
              token sampler beside its eager port), the reference algorithm in bf16 on this GPU (gpu_eager_baseline) and
              the Python-driven decode_n_tokens loop

N > 1 (torchrun, one rank per GPU): every rank decodes its own 64-sample shard of a 64*N global batch (weak scaling),
no collective inside the loop, one NCCL all_gather of the int32 code indices per step inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's debug lines belong on stderr

# stdout carries ONE JSON line: keep the real stdout for it and point fd 1 at stderr for everything else that might write
# there from native code (NCCL prints its version banner to fd 1 whenever NCCL_DEBUG is set in the environment).
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "video-llamagen_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

MODEL, BLOCK, B_LOCAL, NEW_TOKENS = "GPT-L", 576, 64, 576
CFG_SCALE, TOP_K, TOP_P, TEMP = 4.0, 2000, 1.0, 1.0
METRIC = "generated visual tokens/sec, GPT-L c2i-384 CFG decode"
UNIT = "tokens/s"


def workload_config(n_gpus):
    return {"workload": "cfg-2: GPT-L c2i 384px (24x24=576 tokens) bf16, batch 64 with CFG 4.0 (128 rows), top-k 2000",
            "per_gpu_batch": B_LOCAL, "global_batch": B_LOCAL * n_gpus, "new_tokens": NEW_TOKENS,
            "parallelism": f"sample-sharded dp{n_gpus}, all_gather of code indices",
            "l2": "inputs larger than L2 (7.3 GB KV cache + 0.65 GB weights streamed per step; no flush needed)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------- algorithmic bytes
def algorithmic_bytes_generate(cfg, rows, T, N, V):
    d, L, hid = cfg.dim, cfg.n_layer, cfg.hidden_dim
    W = (L * (4 * d * d + 3 * d * hid) + V * d) * 2
    total = 0
    for p in range(T - 1, T + N - 1):                 # position of the input token of each of the N forward passes
        total += W + rows * 2 * d * 2 * L * (p + 1) + rows * 2 * d * 2 * L + rows * V * 4 * 2
    return total


def attention_kernel_roofline(dev, rows, H, S_max, n_pos, peak_gbs, src):
    """Time the attention kernel alone (CUDA events on its stream) over the position mix of one generate():
    launches at pos = 0..n_pos-1, cycling over several distinct cache regions so no launch finds its K/V in L2."""
    import torch
    from vlg import _lib as L
    D = 64
    regions = 6                                            # 6 x 306 MB of K+V >> 126 MB of L2
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(rows, H, D, device=dev, generator=g).bfloat16()
    ks = [torch.randn(rows, H, S_max, D, device=dev, generator=g).bfloat16() for _ in range(regions)]
    vs = [torch.randn(rows, H, S_max, D, device=dev, generator=g).bfloat16() for _ in range(regions)]
    out = torch.empty(rows, H * D, device=dev, dtype=torch.bfloat16)
    lib = L.lib()
    st = torch.cuda.Stream(device=dev)
    positions = list(range(0, n_pos))
    def launch(p, r):
        L.check(lib.vlg_attn_decode(L.ptr(q), L.ptr(ks[r]), L.ptr(vs[r]), L.ptr(out), rows, H, D, S_max, p, None, 0,
                                    L.VLG_BF16, 2, None, 0, L.stream()))
    with torch.cuda.stream(st):
        for p in (0, n_pos // 2, n_pos - 1):               # warm-up
            launch(p, 0)
        st.synchronize()
        # The launches are replayed from a CUDA graph (as in the real decode step) so that the host launch rate of the
        # Python binding does not leak into the device time of the short, small-position launches.
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for i, p in enumerate(positions):
                launch(p, i % regions)
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.synchronize()
        e0.record(st)
        graph.replay()
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1)
    traffic = None                                          # DRAM bytes per launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r1_attention_traffic.json")
    if os.path.exists(tpath) and (rows, H, n_pos) == (128, 16, 576):
        t = json.load(open(tpath))
        traffic = t["dram_bytes_read_per_launch"] + t["dram_bytes_write_per_launch"]
    bytes_total = sum(rows * H * (p + 1) * D * 2 * 2 + rows * H * D * 2 * 2 for p in positions)   # K,V prefix + q + out
    achieved = bytes_total / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "attn_decode_tma_kernel", "achieved": round(achieved, 1), "peak": peak_gbs,
            "unit": "GB/s", "frac": round(achieved / peak_gbs, 4), "traffic": traffic, "peak_source": src,
            "traffic_note": "ncu dram__bytes_read+write of one launch at cache position 287 (= mean position of the timed mix)",
            "launches_timed": len(positions), "avg_launch_us": round(ms * 1e3 / len(positions), 2),
            "algorithmic_bytes_per_launch_avg": bytes_total // len(positions)}


def layer_tail_roofline(model, dev, rows, new_tokens, peak_gbs, src):
    """Time the layer-tail kernel alone: the 1 + n_layer launches of one decode position back to back on the real weights
    (650 MB streamed per replay >> L2), attention skipped (vlg_debug_trunk_without_attention), replayed from a CUDA graph."""
    import torch
    from vlg import _lib as L
    eng = model.engine(rows, 1 + new_tokens)
    eng.begin(rows)
    lib = L.lib()
    st = torch.cuda.Stream(device=dev)
    cfg = model.config
    hid = model.layers[0].feed_forward.w1.weight.shape[0]
    with torch.cuda.stream(st):
        L.check(lib.vlg_debug_trunk_without_attention(eng.ctx, 100, L.stream()))
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        reps = 8
        with torch.cuda.graph(graph, stream=st):
            for i in range(reps):
                L.check(lib.vlg_debug_trunk_without_attention(eng.ctx, 100 + i, L.stream()))
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.synchronize()
        e0.record(st)
        graph.replay()
        e1.record(st)
        st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * (cfg.n_layer + 1))
    w_bytes = (4 * cfg.dim * cfg.dim + 3 * cfg.dim * hid) * 2                 # wo + w1|w3 + w2 + wqkv(next layer), bf16
    act_bytes = rows * (4 * cfg.dim + 2 * hid) * 2                            # attn in, h in/out, xn, ffn, q/k/v out (order of)
    achieved = (w_bytes + act_bytes) / (us * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_lt_ncu_full_step287_metrics.json")
    if os.path.exists(tpath) and (rows, cfg.dim) == (128, 1024):
        full = [e for e in json.load(open(tpath)) if "layer_tail" in e["kernel"]][1:]
        if full:
            traffic = int(sum(float(e["dram__bytes_read.sum"].split()[0]) + float(e["dram__bytes_write.sum"].split()[0])
                              for e in full) / len(full) * 1e6)
    return {"bound": "hbm", "kernel": "layer_tail_kernel", "achieved": round(achieved, 1), "peak": peak_gbs, "unit": "GB/s",
            "frac": round(achieved / peak_gbs, 4), "traffic": traffic, "peak_source": src, "avg_launch_us": round(us, 2),
            "launches_timed": reps * (cfg.n_layer + 1), "algorithmic_bytes_per_launch": w_bytes + act_bytes,
            "note": "one launch per layer for everything between two attentions; streams the layer's weights once, but is "
                    "paced by the SM's L2 port (CTA pairs halve the activation bytes through it) and by six dependent exchanges through L2 with five grid barriers, not by HBM (DESIGN.md 4.1)"}



# --------------------------------------------------------------------------- the other BASELINE configs (N = 1 extras)
ROOF_TOK_S = {"cfg3": 25.4e3, "cfg4_b8": 5.59e3}     # HBM roofline tok/s per GPU at the measured 6 544 GB/s


def _timed_ms(fn, reps):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def extra_cfg3(dev, peak_scale):
    """cfg-3 shard: GPT-XXL c2i 384 px, 32 samples (64 cache rows) per GPU, 576 tokens, CFG 4.0, top-k 2000."""
    import torch
    from vlg.gpt import GPT_models
    from vlg.generate import generate
    torch.manual_seed(0)
    with torch.device(dev):
        m = GPT_models["GPT-XXL"](block_size=576, cls_token_num=1, model_type="c2i").eval()
        torch.nn.init.normal_(m.output.weight, std=0.02)
    m = m.to(dev, torch.bfloat16)
    cond = torch.randint(0, 1000, (32,), generator=torch.Generator().manual_seed(1234)).to(dev)
    fn = lambda: generate(m, cond, 576, cfg_scale=CFG_SCALE, temperature=TEMP, top_k=TOP_K, top_p=TOP_P, sample_logits=True)
    fn()
    ms, out = _timed_ms(fn, 2)
    assert out.shape == (32, 576) and int(out.min()) >= 0 and int(out.max()) < 16384
    v = 32 * 576 / (ms * 1e-3)
    return {"workload": "cfg-3 shard: GPT-XXL c2i 384px, 32 samples (64 rows) per GPU, 576 tokens, bf16, CFG 4.0, top-k 2000",
            "value": round(v, 1), "unit": "tokens/s per GPU", "ms_per_generate": round(ms, 1),
            "hbm_roofline_frac": round(v / (ROOF_TOK_S["cfg3"] * peak_scale), 4)}


def extra_cfg4(dev, peak_scale, B=8):
    """cfg-4: GPT-XL t2v, 5 x 32 x 32 = 5120 latents, 120-token masked caption prefix, CFG 7.5."""
    import torch
    from vlg import gpt_video
    from vlg.generate_video import generate as generate_video
    torch.manual_seed(0)
    with torch.device(dev):
        m = gpt_video.GPT_models["GPT-XL"](block_size=1024, cls_token_num=120, model_type="t2v", vae_embed_dim=8,
                                           num_frames=17, t_downsample_size=4, shuffle_video_tokens=False).eval()
        for n, p in m.named_parameters():
            if "vae_latent_adapter" in n:
                torch.nn.init.normal_(p, std=0.05)
    m = m.to(dev, torch.bfloat16)
    g = torch.Generator().manual_seed(1234)
    cond = torch.randn(B, 120, 2048, generator=g)
    masks = torch.zeros(B, 120)
    for b in range(B):
        masks[b, 120 - int(torch.randint(8, 121, (1,), generator=g)):] = 1
    cond = (cond * masks[:, :, None]).to(dev, torch.bfloat16)
    masks = masks.to(dev)
    generate_video(m, cond, 64, masks, cfg_scale=7.5)                    # warm-up of every kernel on a short run
    ms, out = _timed_ms(lambda: generate_video(m, cond, 5120, masks, cfg_scale=7.5), 1)
    assert out.shape == (B, 5120, 8) and bool(torch.isfinite(out.float()).all())
    v = B * 5120 / (ms * 1e-3)
    return {"workload": f"cfg-4: GPT-XL t2v, B={B} ({2 * B} rows), 5120 latents, 120-token masked prefix, bf16, CFG 7.5",
            "value": round(v, 1), "unit": "latents/s", "ms_per_generate": round(ms, 1),
            "hbm_roofline_frac": round(v / (ROOF_TOK_S["cfg4_b8"] * peak_scale), 4) if B == 8 else None}


def extra_cfg5(dev):
    """cfg-5: VQ index search, 1 048 576 latents x 16384 codes x 8 dims, fp32, index-exact.  Bound: the FP32 FMA pipe
    (SURVEY.md 8d); its peak is measured here with a register-resident FMA loop (scalar and packed)."""
    import ctypes as C
    import torch
    from vlg import _lib as L
    from vlg.vq_model import VectorQuantizer
    torch.manual_seed(0)
    vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval().to(dev)
    z = torch.randn(1024, 8, 32, 32, device=dev)
    for _ in range(3):
        vq(z)
    batches = sorted(_timed_ms(lambda: vq(z), 10)[0] for _ in range(3))       # three batches of ten calls: the median batch
    ms = batches[1]
    out = torch.zeros(4, device=dev)
    peak = {}
    for variant, name in ((0, "ffma"), (1, "ffma2")):
        fl = C.c_longlong()
        f = lambda: L.check(L.lib().vlg_debug_fp32_fma_peak(variant, 4096, L.ptr(out), C.byref(fl), L.stream()))
        f()
        pms, _ = _timed_ms(f, 5)
        peak[name] = round(fl.value / pms / 1e9, 2)
    n = 1 << 20
    tf = 2 * n * 16384 * 8 / ms / 1e9
    return {"workload": "cfg-5: VQ argmin, 1,048,576 latents x 16384 codes x 8 dims, fp32, index-exact", "value": round(n / ms * 1e3),
            "unit": "latents/s", "ms": round(ms, 3), "ms_batches": [round(b, 3) for b in batches], "roofline": {"bound": "fp32_fma", "achieved": round(tf, 2),
            "peak": max(peak.values()), "unit": "TFLOP/s", "frac": round(tf / max(peak.values()), 4),
            "peak_source": "measured here: register-resident FMA loop", "peaks": peak,
            "algorithmic_flops": 2 * n * 16384 * 8}}


def extra_diff_head(dev, rows=1):
    """Diffusion-head token sampler in the configuration the reference ships (GPT-1B trunk: z of 2048,
    head width 1024, depth 3, 100 sampling steps, batch 1, no guidance; scripts/sample/sample_t2v_diff.bash): one
    DiffLoss.sample per generated token.  Latency-bound (a chain of 700 dependent layers per token); reported beside the
    oracle port of the same call through torch library kernels on the same GPU."""
    import time
    import torch
    from vlg.diffloss import DiffLoss
    from oracle import diffusion as D
    torch.manual_seed(0)
    head = DiffLoss(target_channels=8, z_channels=2048, depth=3, width=1024, num_sampling_steps="100").eval()
    for m in head.modules():
        if isinstance(m, torch.nn.Linear):
            m.weight.data.normal_(0, 1.0 / m.weight.shape[1] ** 0.5)
            m.bias.data.normal_(0, 0.05)
    head = head.to(dev)
    z = torch.randn(rows, 2048, device=dev)
    for _ in range(3):
        head.sample(z, 1.0, 1.0, seed=1)
    ms, _ = _timed_ms(lambda: head.sample(z, 1.0, 1.0, seed=2), 20)
    head.engine(rows).check()
    w = dict(head.state_dict())
    noise, step_noise = torch.randn(rows, 8, device=dev), torch.randn(100, rows, 8, device=dev)
    D.p_sample_loop(w, z, noise, step_noise)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D.p_sample_loop(w, z, noise, step_noise)
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t0) * 1e3
    # the same call on this host's cores (the reference's own CPU path is this arithmetic in torch): ONE token, bounded
    cpu_w = {k: v.detach().cpu() for k, v in w.items()}
    t0 = time.perf_counter()
    D.p_sample_loop(cpu_w, z.cpu(), noise.cpu(), step_noise.cpu())
    cpu_ms = (time.perf_counter() - t0) * 1e3
    layer_bytes = 4 * (2 * 3 * 1024 * 1024 + 16 * 1024 + 8 * 1024)       # weights one diffusion step touches (shared memory)
    return {"workload": f"diffusion head: DiffLoss.sample, {rows} row(s), z 2048 -> latent 8, width 1024, depth 3, 100 steps, fp32",
            "value": round(rows / ms * 1e3, 1), "unit": "latents/s", "ms_per_token": round(ms, 3),
            "us_per_diffusion_step": round(ms * 10, 2), "grid_barriers_per_step": 7,
            "bound": "latency: 7 dependent grid-wide exchanges per diffusion step, weights resident in shared memory "
                     f"({layer_bytes * 100 / 1e6:.0f} MB of weight reads per token never leave the SMs)",
            "gpu_eager_port": {"ms_per_token": round(eager_ms, 1), "kind": "oracle port on cuda (torch library kernels, fp32)"},
            "cpu_baseline": {"ms_per_token": round(cpu_ms, 1), "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "one token (100 sampling steps) on the host cores, fp32"}}


def _conv_flops(module, x_shape):
    """2 * MACs of every Conv2d of `module` for one forward at input x_shape, counted with forward hooks on meta tensors."""
    import torch
    total = [0]

    def hook(m, inp, out):
        total[0] += 2 * out.numel() * m.in_channels * m.kernel_size[0] * m.kernel_size[1] // m.groups
    hs = [m.register_forward_hook(hook) for m in module.modules() if isinstance(m, torch.nn.Conv2d)]
    try:
        with torch.no_grad():
            module(torch.zeros(x_shape, device="meta"))
    finally:
        for h in hs:
            h.remove()
    return total[0]


def extra_decode_code(dev, pk, n_images=64, grid=24):
    """The call right after generate() in every sampler (sample_t2i.py:134): code indices -> images through
    VQModel.decode_code.  The codebook gather + post_quant_conv is one vlg kernel; the decoder CNN is library convolutions
    (cuDNN through torch, TF32 allowed as the reference scripts set it, sample_t2i.py:2-4) -- tensor-bound, so the fraction is
    against the measured dense bf16 GEMM peak halved (TF32 runs at half the bf16 rate)."""
    import copy
    import torch
    from vlg.vq_model import VQ_models
    torch.manual_seed(0)
    vq = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).eval()
    flops = _conv_flops(copy.deepcopy(vq.decoder).to("meta"), (1, 256, grid, grid))
    vq = vq.to(dev)
    codes = torch.randint(0, 16384, (n_images * grid * grid,), device=dev)
    shape = (n_images, 8, grid, grid)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        with torch.no_grad():
            for _ in range(2):
                img = vq.decode_code(codes, shape)
            ms, _ = _timed_ms(lambda: vq.decode_code(codes, shape), 5)
            ms_front, _ = _timed_ms(lambda: vq.quantize.decode_front(codes, shape, vq.post_quant_conv.weight, vq.post_quant_conv.bias), 20)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    tf = flops * n_images / ms / 1e9
    peak = pk.get("bf16_tflops", 1603.9) / 2
    return {"workload": f"decode_code: {n_images} x {grid}x{grid} codes -> {n_images} x 3 x {grid * 16}x{grid * 16} images (VQ-16 decoder, fp32 / TF32 convolutions)",
            "value": round(n_images / ms * 1e3, 1), "unit": "images/s", "ms": round(ms, 2), "image_shape": list(img.shape),
            "fused_gather_post_quant_conv_ms": round(ms_front, 4),
            "roofline": {"bound": "tensor", "achieved": round(tf, 1), "peak": round(peak, 1), "unit": "TFLOP/s", "frac": round(tf / peak, 4),
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32)", "conv_flops_per_image": flops,
                         "note": "library convolutions (cuDNN): outside the hand-written hot path, reported because it is the wall-clock tail after generate()"}}


def extra_gpu_eager(dev, decode_steps):
    """The reference algorithm (oracle port: torch library kernels, the reference's op sequence incl. its full-cache
    attention) in bf16 on this GPU, on the sample the CPU leg runs: the like-for-like successor of the A100 eager column of
    BASELINE.md.  Test infrastructure timed as a baseline, never on the product path."""
    import torch
    from oracle import ar_decode as O
    geo = O.Geometry(block_size=BLOCK, **O.NAMED[MODEL])
    w = {k: v.to(dev, torch.bfloat16) for k, v in O.init_weights(geo, seed=0).items()}
    cond = torch.randint(0, 1000, (B_LOCAL,), generator=torch.Generator().manual_seed(1234)).to(dev)
    rows = torch.cat([cond, torch.ones_like(cond) * geo.num_classes])

    def run(n_steps):
        st = O.DecodeState(geo, 2 * B_LOCAL, 1 + NEW_TOKENS, torch.bfloat16, dev)
        logits = O.forward_logits(geo, w, st, cond=rows, input_pos=torch.arange(1, device=dev))
        tok = O.sample(O.cfg_mix(logits, CFG_SCALE), temperature=TEMP, top_k=TOP_K, top_p=TOP_P)[0]
        for i in range(n_steps):
            logits = O.forward_logits(geo, w, st, idx=torch.cat([tok, tok]).view(-1, 1), input_pos=torch.tensor([1 + i], device=dev))
            tok = O.sample(O.cfg_mix(logits, CFG_SCALE), temperature=TEMP, top_k=TOP_K, top_p=TOP_P)[0]
        return tok
    with torch.no_grad():
        run(2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(decode_steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    return {"value": round(B_LOCAL * (decode_steps + 1) / dt, 1), "unit": UNIT, "kind": "port on cuda (torch library kernels, bf16)",
            "sample": f"cfg-2 rows: prefill + {decode_steps} decode steps of the 576 ({dt * 1e3:.0f} ms, cache alloc included as in "
                      "the reference's generate()); per-step cost position-independent (attends over all S_max slots)"}


def extra_python_loop(model, cond_dev, n_tokens=64):
    """The reference's own call pattern -- prefill() + decode_n_tokens() driven from Python, one C-ABI call per position
    (generate.py:170-178) -- next to the in-library graph loop."""
    import torch
    from vlg.generate import prefill, decode_n_tokens
    dev = cond_dev.device
    B = cond_dev.shape[0]
    kw = dict(temperature=TEMP, top_k=TOP_K, top_p=TOP_P, sample_logits=True)
    rows = torch.cat([cond_dev, torch.ones_like(cond_dev) * 1000])

    def run():
        model.setup_caches(2 * B, 1 + n_tokens, None)
        tok = prefill(model, rows, torch.arange(1, device=dev), CFG_SCALE, **kw)
        toks, _ = decode_n_tokens(model, tok, torch.tensor([1], device=dev, dtype=torch.int), n_tokens - 1, CFG_SCALE, -1, **kw)
        return toks
    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": round(B * n_tokens / dt, 1), "unit": UNIT, "tokens": n_tokens,
            "note": "Python-driven prefill + decode_n_tokens over the first 64 positions (per-position C-ABI calls, probabilities "
                    "materialised every step as the reference does); generate() replays a CUDA graph instead"}


# --------------------------------------------------------------------------- CPU reference arm (oracle port)
def cpu_reference_sample(decode_steps, threads=None):
    """Reference algorithm (oracle/ar_decode.py restatement of generate.py + gpt.py) on the host cores: cfg-2 rows and
    geometry, fp32 on CPU, prefill + `decode_steps` decode steps.  Its per-step cost is position-independent because the
    reference attends over all S_max slots and copies both caches every layer (gpt.py:230-237)."""
    import torch
    from oracle import ar_decode as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    geo = O.Geometry(block_size=BLOCK, **O.NAMED[MODEL])
    w = O.init_weights(geo, seed=0)
    cond = torch.randint(0, 1000, (B_LOCAL,), generator=torch.Generator().manual_seed(1234))
    n = decode_steps + 1
    # generate() allocates caches for the full 576-token sequence (as the reference does); run only n tokens of it
    t0 = time.perf_counter()
    with torch.no_grad():
        rows = torch.cat([cond, torch.ones_like(cond) * geo.num_classes])
        st = O.DecodeState(geo, 2 * B_LOCAL, 1 + NEW_TOKENS, torch.float32)
        t_alloc = time.perf_counter() - t0
        t0 = time.perf_counter()
        logits = O.forward_logits(geo, w, st, cond=rows, input_pos=torch.arange(1))
        tok = O.sample(O.cfg_mix(logits, CFG_SCALE), temperature=TEMP, top_k=TOP_K, top_p=TOP_P)[0]
        for i in range(decode_steps):
            logits = O.forward_logits(geo, w, st, idx=torch.cat([tok, tok]).view(-1, 1), input_pos=torch.tensor([1 + i]))
            tok = O.sample(O.cfg_mix(logits, CFG_SCALE), temperature=TEMP, top_k=TOP_K, top_p=TOP_P)[0]
    dt = time.perf_counter() - t0
    return {"value": round(B_LOCAL * n / dt, 2), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"cfg-2 rows on CPU fp32: prefill + {decode_steps} decode steps of the 576 "
                      f"({dt:.1f} s; cache alloc {t_alloc:.1f} s excluded), per-step cost position-independent"}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, last = [], None
    for i in range(args.warmup + args.steps):
        last, dt = cpu_reference_sample(args.ref_decode_steps)
        if i >= args.warmup:
            times.append(dt)
    tok = B_LOCAL * (args.ref_decode_steps + 1)
    value = tok / statistics.mean(times)
    last["value"] = round(value, 2)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(statistics.mean(times) * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "cpu_baseline": last,
            "e2e": {"value": round(value, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# --------------------------------------------------------------------------- GPU arm
def run_vlg(args):
    import torch
    import torch.distributed as dist
    from vlg import _lib as L
    from vlg.gpt import GPT_models
    from vlg.generate import generate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the vlg path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}"

    torch.manual_seed(0)
    model = GPT_models[MODEL](block_size=BLOCK, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(model.output.weight, std=0.02)      # zero-initialised in the reference (gpt.py:307)
    model = model.to(dev, torch.bfloat16)
    gen = torch.Generator().manual_seed(1234)
    cond_all = torch.randint(0, 1000, (B_LOCAL * world,), generator=gen)
    cond_host = cond_all[rank * B_LOCAL:(rank + 1) * B_LOCAL].contiguous().pin_memory()
    cond_dev = cond_host.to(dev)
    tokens_host = torch.empty(B_LOCAL, NEW_TOKENS, dtype=torch.int32).pin_memory()
    gathered = torch.empty(world * B_LOCAL, NEW_TOKENS, dtype=torch.int32, device=dev) if world > 1 else None
    kw = dict(cfg_scale=CFG_SCALE, temperature=TEMP, top_k=TOP_K, top_p=TOP_P, sample_logits=True)
    stream = torch.cuda.Stream(device=dev)
    lib = L.lib()

    def one_step(seed):
        toks = generate(model, cond_dev, NEW_TOKENS, seed=1000 * seed + rank, **kw)
        if world > 1:
            dist.all_gather_into_tensor(gathered, toks)
        return toks

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        fence()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(n):
            fn(i)
        e1.record(stream)
        fence()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.cuda.stream(stream):
        for i in range(args.warmup):
            toks = one_step(100 + i)
        torch.cuda.synchronize()
        if args.warmup:
            assert int(toks.min()) >= 0 and int(toks.max()) < 16384
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = lib.vlg_launch_count()
        ms = timed(one_step, args.steps)
        launches = lib.vlg_launch_count() - launches0

        if args.skip_extras:
            if rank == 0:
                sampler.stop()
                emit({"metric": METRIC, "value": round(B_LOCAL * NEW_TOKENS * world * args.steps / (ms * 1e-3), 1),
                      "unit": UNIT, "ms_per_step": round(ms / args.steps, 2), "new_tokens": NEW_TOKENS,
                      "gpu_launches": int(launches), "mode": "profiling run (skip-extras): not a bench value"})
            return
        # e2e: host labels -> host tokens through the C-ABI call (copies + sync inside)
        eng = model._engine

        def e2e_step(i):
            eng.generate_host(cond_host, tokens_host, NEW_TOKENS, CFG_SCALE, -1, TEMP, TOP_K, TOP_P, True, 2000 + i)
            if world > 1:
                dist.all_gather_into_tensor(gathered, tokens_host.to(dev, non_blocking=True))
        e2e_step(0)
        fence()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i + 1)
        fence()
        e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_s.item())
        clocks = sampler.stop() if rank == 0 else None

    tokens_per_step = B_LOCAL * NEW_TOKENS * world
    value = tokens_per_step * args.steps / (ms * 1e-3)
    if rank == 0:
        pk, src = peaks()
        cfg = model.config
        geo_hidden = model.layers[0].feed_forward.w1.weight.shape[0]

        class G:  # tiny adapter for algorithmic_bytes_generate
            dim, n_layer, hidden_dim = cfg.dim, cfg.n_layer, geo_hidden
        step_bytes = algorithmic_bytes_generate(G, 2 * B_LOCAL, 1, NEW_TOKENS, cfg.vocab_size)
        roof_a = attention_kernel_roofline(dev, 2 * B_LOCAL, cfg.n_head, 584, NEW_TOKENS, pk["hbm_gbs"], src)
        roof_l = layer_tail_roofline(model, dev, 2 * B_LOCAL, NEW_TOKENS, pk["hbm_gbs"], src)
        step_ms = ms / args.steps
        step_gbs = step_bytes / (step_ms * 1e-3) / 1e9
        # share of one generate(): launches per generate x the kernel's average duration timed alone
        roof_a["share_of_step"] = round(cfg.n_layer * NEW_TOKENS * roof_a["avg_launch_us"] * 1e-3 / step_ms, 4)
        roof_l["share_of_step"] = round((cfg.n_layer + 1) * NEW_TOKENS * roof_l["avg_launch_us"] * 1e-3 / step_ms, 4)
        whole = {"bound": "hbm", "achieved": round(step_gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                 "frac": round(step_gbs / pk["hbm_gbs"], 4), "algorithmic_bytes_per_step": step_bytes, "peak_source": src,
                 "note": "SURVEY.md 8d bytes of one generate() / its measured time: the figure the >= 0.60 target is quoted on"}
        dominant, other = (roof_l, roof_a) if roof_l["share_of_step"] >= roof_a["share_of_step"] else (roof_a, roof_l)
        roof = dict(dominant)
        roof["whole_step_achieved_gbs"] = whole["achieved"]
        roof["whole_step_frac"] = whole["frac"]
        cpu = None
        extras = {}
        if world == 1 and not args.no_cpu_baseline:
            cpu, _ = cpu_reference_sample(args.ref_decode_steps)
        if world == 1 and not args.no_extras:
            peak_scale = pk["hbm_gbs"] / 6544.3
            for name, fn in (("python_driven_loop", lambda: extra_python_loop(model, cond_dev)),
                             ("gpu_eager_baseline", lambda: extra_gpu_eager(dev, args.ref_decode_steps)),
                             ("cfg5_vq", lambda: extra_cfg5(dev)),
                             ("decode_code", lambda: extra_decode_code(dev, pk)),
                             ("diff_head", lambda: extra_diff_head(dev)),
                             ("cfg3", lambda: extra_cfg3(dev, peak_scale)),
                             ("cfg4_b8", lambda: extra_cfg4(dev, peak_scale, 8))):
                try:
                    if name == "cfg3":                    # the cfg-2 model and its 7.3 GB cache are no longer needed
                        model._engine = None
                        torch.cuda.empty_cache()
                    extras[name] = fn()
                except Exception as e:                    # an extra never costs the headline line
                    extras[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                torch.cuda.empty_cache()
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 2), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks,
                "e2e": {"value": round(tokens_per_step * args.steps / e2e_s, 1), "unit": UNIT,
                        "h2d_bytes_per_step": B_LOCAL * 8 * world, "d2h_bytes_per_step": B_LOCAL * NEW_TOKENS * 4 * world},
                "gpu_launches": int(launches) * world, "roofline": roof,
                "roofline_" + ("attention" if other is roof_a else "layer_tail"): other, "roofline_whole_step": whole,
                "cpu_baseline": cpu}
        line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="vlg", choices=["vlg", "reference"])
    ap.add_argument("--ref-decode-steps", type=int, default=5, help="decode steps per CPU-reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg-3 / cfg-4 / cfg-5 / GPU-eager / Python-loop legs")
    ap.add_argument("--new-tokens", type=int, default=NEW_TOKENS, help="tokens per sample (576 = the BASELINE workload)")
    ap.add_argument("--skip-extras", action="store_true",
                    help="profiling mode: only the device-resident generate() loop (no e2e / roofline / CPU legs)")
    args = ap.parse_args()
    globals()["NEW_TOKENS"] = args.new_tokens
    if args.impl == "reference":
        run_reference(args)
    else:
        run_vlg(args)


if __name__ == "__main__":
    main()
