"""Drop-in for the inference surface of the reference's ``autoregressive/models/diffloss.py`` (the diffusion head of
``gpt_video_diff``).

``DiffLoss(target_channels, z_channels, depth, width, num_sampling_steps)`` keeps the reference's module tree -- so
``state_dict`` keys (``net.time_embed.mlp.0.weight`` ... ``net.final_layer.linear.bias``) and the weights a given seed
produces are the reference's (diffloss.py:151-210) -- and ``sample(z, temperature, cfg)`` (diffloss.py:35-52) runs the
whole ``num_sampling_steps``-long chain in ONE kernel launch of libvlg.so (``vlg_diff_sample``, csrc/diff_head.cu)
instead of ~4000 eager launches.  fp32 only, like the reference's shipped sampler (``--precision none``).

Host side (as in the reference, where ``create_diffusion`` builds float64 numpy tables, diffusion/__init__.py:11-47):
the cosine schedule, its respacing and the six per-step coefficients are derived here in float64 and handed to the
library as fp32 tables.  The training loss (``forward``) is out of scope and raises.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

_TABLE_ORDER = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
                "posterior_log_variance_clipped", "log_betas")


def sampling_schedule(num_sampling_steps="100", diffusion_steps=1000):
    """The constants of ``create_diffusion(timestep_respacing=num_sampling_steps, noise_schedule="cosine")``.

    Cosine alpha-bar schedule with betas capped at 0.999 (gaussian_diffusion.py:115-150); the kept timesteps are evenly
    strided inside each comma-separated section (respace.py:12-60); the respaced betas are re-derived from the kept
    cumulative products (respace.py:71-87); coefficient tables as gaussian_diffusion.py:172-201.  Returns
    (timestep_map int32 [S], tables float64 [6, S] in the order vlg_diff_create expects)."""
    spec = str(num_sampling_steps)
    if spec.startswith("ddim"):
        raise NotImplementedError("ddimN respacing is not used by DiffLoss")
    counts = [int(c) for c in spec.split(",")]
    # cosine schedule
    grid = np.arange(diffusion_steps + 1, dtype=np.float64) / diffusion_steps
    bar = np.array([math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2 for t in grid])
    base_betas = np.minimum(1.0 - bar[1:] / bar[:-1], 0.999)
    base_cumprod = np.cumprod(1.0 - base_betas)
    # kept timesteps, section by section: a running float position advanced by the section's stride and rounded
    kept, start = set(), 0
    for sec, count in enumerate(counts):
        size = diffusion_steps // len(counts) + (1 if sec < diffusion_steps % len(counts) else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.add(start + round(pos))
            pos += stride
        start += size
    tmap = np.array(sorted(kept), dtype=np.int64)
    ac = base_cumprod[tmap]
    ac_prev = np.concatenate([[1.0], ac[:-1]])
    betas = 1.0 - ac / ac_prev
    alphas = 1.0 - betas
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    tables = np.stack([
        np.sqrt(1.0 / ac),
        np.sqrt(1.0 / ac - 1.0),
        betas * np.sqrt(ac_prev) / (1.0 - ac),
        (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        np.log(np.concatenate([post_var[1:2], post_var[1:]])),
        np.log(betas),
    ])
    return tmap.astype(np.int32), tables


class TimestepEmbedder(nn.Module):
    """diffloss.py:59-97 (parameter holder; the embedding table of the sampling timesteps is computed by the library)."""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size


class ResBlock(nn.Module):
    """diffloss.py:100-127."""

    def __init__(self, channels):
        super().__init__()
        self.channels = channels
        self.in_ln = nn.LayerNorm(channels, eps=1e-6)
        self.mlp = nn.Sequential(nn.Linear(channels, channels, bias=True), nn.SiLU(), nn.Linear(channels, channels, bias=True))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(channels, 3 * channels, bias=True))


class FinalLayer(nn.Module):
    """diffloss.py:130-149."""

    def __init__(self, model_channels, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(model_channels, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(model_channels, out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(model_channels, 2 * model_channels, bias=True))


class SimpleMLPAdaLN(nn.Module):
    """diffloss.py:151-210: same sub-modules in the same order, same initialisation."""

    def __init__(self, in_channels, model_channels, out_channels, z_channels, num_res_blocks, grad_checkpointing=False):
        super().__init__()
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.grad_checkpointing = grad_checkpointing
        self.time_embed = TimestepEmbedder(model_channels)
        self.cond_embed = nn.Linear(z_channels, model_channels)
        self.input_proj = nn.Linear(in_channels, model_channels)
        self.res_blocks = nn.ModuleList([ResBlock(model_channels) for _ in range(num_res_blocks)])
        self.final_layer = FinalLayer(model_channels, out_channels)
        self.initialize_weights()

    def initialize_weights(self):
        def _basic_init(module):
            if isinstance(module, nn.Linear):
                torch.nn.init.xavier_uniform_(module.weight)
                if module.bias is not None:
                    nn.init.constant_(module.bias, 0)
        self.apply(_basic_init)
        nn.init.normal_(self.time_embed.mlp[0].weight, std=0.02)
        nn.init.normal_(self.time_embed.mlp[2].weight, std=0.02)
        for block in self.res_blocks:
            nn.init.constant_(block.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(block.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    def forward(self, x, t, c):
        raise NotImplementedError("vlg runs the network inside DiffLoss.sample (one kernel for the whole chain); a single "
                                  "forward is only needed by the training loss, which is out of scope (DESIGN.md)")


class DiffEngine:
    """Owns one ``vlg_diff`` handle: borrowed fp32 weights, workspace, schedule tables."""

    MAX_ROWS = 16

    def __init__(self, head, max_rows):
        net = head.net
        p = net.input_proj.weight
        if not p.is_cuda:
            raise L.VlgError("DiffLoss is on CPU: the vlg diffusion-head sampler needs a B200 (no CPU fallback)")
        if p.dtype != torch.float32:
            raise L.VlgError(f"DiffLoss weights are {p.dtype}: the diffusion head runs in float32 (the reference's sampler "
                             "runs with --precision none)")
        self.device = p.device
        self.max_rows = int(max_rows)
        self.keep = []

        def dev(t):
            t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        W, Cc = net.model_channels, net.in_channels
        Z = net.cond_embed.weight.shape[1]
        depth = len(net.res_blocks)
        tmap, tables = sampling_schedule(head.num_sampling_steps)
        self.steps = int(tmap.shape[0])
        self.geo = L.DiffGeometry(Cc, W, Z, depth, self.steps)
        self.C, self.Z = Cc, Z
        blocks = (L.DiffBlock * depth)()
        for i, b in enumerate(net.res_blocks):
            blocks[i] = L.DiffBlock(dev(b.in_ln.weight), dev(b.in_ln.bias), dev(b.mlp[0].weight), dev(b.mlp[0].bias),
                                    dev(b.mlp[2].weight), dev(b.mlp[2].bias), dev(b.adaLN_modulation[1].weight),
                                    dev(b.adaLN_modulation[1].bias))
        self.blocks = blocks
        w = L.DiffWeights()
        w.time_w0, w.time_b0 = dev(net.time_embed.mlp[0].weight), dev(net.time_embed.mlp[0].bias)
        w.time_w1, w.time_b1 = dev(net.time_embed.mlp[2].weight), dev(net.time_embed.mlp[2].bias)
        w.cond_w, w.cond_b = dev(net.cond_embed.weight), dev(net.cond_embed.bias)
        w.in_w, w.in_b = dev(net.input_proj.weight), dev(net.input_proj.bias)
        w.blocks = blocks
        w.final_ada_w, w.final_ada_b = dev(net.final_layer.adaLN_modulation[1].weight), dev(net.final_layer.adaLN_modulation[1].bias)
        w.final_w, w.final_b = dev(net.final_layer.linear.weight), dev(net.final_layer.linear.bias)
        self.weights = w
        lib = L.lib()
        nbytes = lib.vlg_diff_workspace_bytes(C.byref(self.geo), self.max_rows)
        if nbytes == 0:
            raise L.VlgError(f"unsupported diffusion-head geometry C={Cc} W={W} depth={depth} steps={self.steps} rows={max_rows}: "
                             + lib.vlg_last_error().decode())
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        tmap_c = np.ascontiguousarray(tmap, dtype=np.int32)
        tab_c = np.ascontiguousarray(tables.astype(np.float32))          # the reference rounds each extracted entry to fp32
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(lib.vlg_diff_create(C.byref(self.geo), C.byref(w), tmap_c.ctypes.data_as(C.c_void_p),
                                        tab_c.ctypes.data_as(C.c_void_p), self.max_rows, C.c_void_p(base), nbytes,
                                        C.byref(self.handle)))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.lib().vlg_diff_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def check(self):
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_diff_check(self.handle, L.stream()))

    def sample(self, z, temperature, cfg, noise=None, step_noise=None, seed=0, offset=0):
        rows = z.shape[0]
        z = z.detach().to(device=self.device, dtype=torch.float32).contiguous()
        assert z.shape == (rows, self.Z), f"z must be [rows, {self.Z}], got {tuple(z.shape)}"
        out = torch.empty(rows, self.C, dtype=torch.float32, device=self.device)
        if noise is not None:
            noise = noise.detach().to(device=self.device, dtype=torch.float32).contiguous()
            assert noise.shape == (rows, self.C)
        if step_noise is not None:
            step_noise = step_noise.detach().to(device=self.device, dtype=torch.float32).contiguous()
            assert step_noise.shape == (self.steps, rows, self.C)
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_diff_sample(self.handle, L.ptr(z), rows, float(temperature), float(cfg), L.ptr(noise),
                                            L.ptr(step_noise), int(seed), int(offset), L.ptr(out), L.stream()))
        return out


class DiffLoss(nn.Module):
    """Diffusion Loss head (diffloss.py:10-52): inference (``sample``) only."""

    def __init__(self, target_channels, z_channels, depth, width, num_sampling_steps, grad_checkpointing=False):
        super().__init__()
        self.in_channels = target_channels
        self.net = SimpleMLPAdaLN(in_channels=target_channels, model_channels=width, out_channels=target_channels * 2,
                                  z_channels=z_channels, num_res_blocks=depth, grad_checkpointing=grad_checkpointing)
        self.num_sampling_steps = str(num_sampling_steps)
        self._engine = None
        self._calls = 0

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, *a, **k):
        self._engine = None
        return super().load_state_dict(state_dict, *a, **k)

    def engine(self, rows):
        e = self._engine
        if e is None or e.max_rows < rows:
            self._engine = None
            self._engine = DiffEngine(self, max(int(rows), 1))
        return self._engine

    def forward(self, target, z, mask=None):
        raise NotImplementedError("vlg implements DiffLoss.sample only (the training loss is out of scope: DESIGN.md)")

    @torch.no_grad()
    def sample(self, z, temperature=1.0, cfg=1.0, *, noise=None, step_noise=None, seed=None):
        """diffloss.py:35-52.  z [rows, z_channels] -> sampled latents [rows, target_channels] (fp32).

        With ``cfg != 1`` the rows are (conditional half, unconditional half) and the start noise is shared by the halves
        (diffloss.py:37-41).  ``noise`` [rows, C] / ``step_noise`` [steps, rows, C] replace the in-kernel Philox draws
        (parity tests feed the reference's own draws); otherwise a seed is taken from torch's CPU generator, so
        ``torch.manual_seed`` makes runs reproducible."""
        L.require_cuda(z, "DiffLoss.sample")
        n, cap = z.shape[0], DiffEngine.MAX_ROWS
        if n > cap:
            if cfg != 1.0:
                raise L.VlgError(f"DiffLoss.sample with guidance handles at most {cap} rows per call, got {n}")
            parts = []
            for i in range(0, n, cap):
                parts.append(self.sample(z[i:i + cap], temperature, cfg, noise=None if noise is None else noise[i:i + cap],
                                         step_noise=None if step_noise is None else step_noise[:, i:i + cap], seed=seed))
            return torch.cat(parts)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
        self._calls += 1
        return self.engine(z.shape[0]).sample(z, temperature, cfg, noise, step_noise, seed, self._calls)
