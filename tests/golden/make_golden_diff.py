"""Generate ``diff_head.npz`` / ``diff_generate.npz`` from the LIVE reference (authoring container only):

    python tests/golden/make_golden_diff.py

The reference's diffusion-head sampler (DiffLoss.sample -> SpacedDiffusion.p_sample_loop over SimpleMLPAdaLN,
diffloss.py:35-52, gaussian_diffusion.py:376-517) is run UNMODIFIED on CPU; two things are patched around it, not in it:
``Tensor.cuda`` becomes the identity (the reference moves its noise and timestep tensors with ``.cuda()``), and
``torch.randn`` / ``torch.randn_like`` record what they return, so that the oracle and the CUDA path can be fed the same
draws.  The schedule tables of the reference's ``gen_diffusion`` object are stored too (the oracle derives its own and is
compared with them).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import _reference  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.set_float32_matmul_precision("highest")


def save(name, **arrays):
    arrays["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


class Recorder:
    """Records every torch.randn / torch.randn_like result while active."""

    def __enter__(self):
        self.draws = []
        self._randn, self._randn_like, self._cuda = torch.randn, torch.randn_like, torch.Tensor.cuda

        def randn(*a, **k):
            t = self._randn(*a, **k)
            self.draws.append(t.clone())
            return t

        def randn_like(x, **k):
            t = self._randn_like(x, **k)
            self.draws.append(t.clone())
            return t

        torch.randn, torch.randn_like = randn, randn_like
        torch.Tensor.cuda = lambda self_, *a, **k: self_
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like, torch.Tensor.cuda = self._randn, self._randn_like, self._cuda
        return False


def trained_like(module, gen):
    """A random-init head has zero adaLN / output layers (diffloss.py:196-210) and would emit zeros: give every Linear and
    LayerNorm parameter a seeded non-trivial value, at the scale gpt_video_diff's own init uses (normal 0.02 weights,
    gpt_video_diff.py:437-444) with biases switched on."""
    for m in module.modules():
        if isinstance(m, torch.nn.Linear):
            m.weight.data = torch.randn(m.weight.shape, generator=gen) * (1.0 / m.weight.shape[1] ** 0.5)
            m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.05
        elif isinstance(m, torch.nn.LayerNorm) and m.elementwise_affine:
            m.weight.data = 1.0 + 0.1 * torch.randn(m.weight.shape, generator=gen)
            m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=gen)


def split_draws(draws, steps):
    """DiffLoss.sample draws the start noise first (diffloss.py:38/43), then p_sample draws once per loop index from
    steps-1 down to 0 (gaussian_diffusion.py:411)."""
    assert len(draws) == 1 + steps
    noise = draws[0]
    step_noise = torch.stack(draws[1:][::-1])          # index i = loop index
    return noise, step_noise


def head_fixture(ref):
    gen = torch.Generator().manual_seed(20)
    C, Z, W, depth = 8, 96, 128, 3
    d = ref.diffloss.DiffLoss(target_channels=C, z_channels=Z, depth=depth, width=W, num_sampling_steps="100").eval()
    trained_like(d, gen)
    out = {("w." + k): v.numpy() for k, v in d.state_dict().items()}
    gd = d.gen_diffusion
    for name in ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                 "posterior_mean_coef2", "posterior_log_variance_clipped"):
        out["tab." + name] = np.asarray(getattr(gd, name))
    out["tab.log_betas"] = np.log(gd.betas)
    out["tab.timestep_map"] = np.asarray(gd.timestep_map, dtype=np.int64)
    # one forward of the net alone (original-scale timestep), then the whole loop without and with guidance
    x = torch.randn(4, C, generator=gen)
    c = torch.randn(4, Z, generator=gen)
    t = torch.tensor([999, 500, 10, 0])
    with torch.no_grad():
        out["net.x"], out["net.c"], out["net.t"] = x.numpy(), c.numpy(), t.numpy()
        out["net.out"] = d.net(x, t, c).numpy()
        out["net.out_cfg"] = d.net.forward_with_cfg(x, t, c, 3.0).numpy()
    for tag, rows, temperature, cfg in (("plain", 3, 1.0, 1.0), ("temp", 2, 0.7, 1.0), ("cfg", 4, 0.9, 3.0)):
        z = torch.randn(rows, Z, generator=gen)
        torch.manual_seed(100 + rows)
        with Recorder() as rec, torch.no_grad():
            sample = d.sample(z, temperature, cfg)
        noise, step_noise = split_draws(rec.draws, 100)
        if cfg != 1.0:                                     # diffloss.py:38-39: half the rows drawn, then duplicated
            noise = torch.cat([noise, noise], dim=0)
        out[f"{tag}.z"], out[f"{tag}.noise"], out[f"{tag}.step_noise"] = z.numpy(), noise.numpy(), step_noise.numpy()
        out[f"{tag}.temperature"], out[f"{tag}.cfg"] = np.float32(temperature), np.float32(cfg)
        out[f"{tag}.sample"] = sample.numpy()
        print(tag, "sample rms", float(sample.pow(2).mean().sqrt()), "absmax", float(sample.abs().max()))
    save("diff_head.npz", **out)


def generate_fixture(ref):
    """generate_video_diff.generate (cfg_scale 1, B = 1) on a tiny gpt_video_diff trunk + head."""
    torch.manual_seed(7)
    M = ref.gpt_video_diff
    args = M.ModelArgs(dim=128, n_layer=2, n_head=2, block_size=16, cls_token_num=6, caption_dim=32, vae_embed_dim=8,
                       num_frames=5, t_downsample_size=4, vocab_size=64, diffloss_w=128, diffloss_d=3, num_sampling_steps="100",
                       model_type="t2v", multiple_of=64)
    model = M.Transformer(args).eval()
    gen = torch.Generator().manual_seed(8)
    trained_like(model.diffloss, gen)
    # lively trunk: the stock 0.02 init leaves the hidden state nearly input-independent
    for n, p in model.named_parameters():
        if n.startswith("layers.") and p.dim() == 2:
            p.data = torch.randn(p.shape, generator=gen) * (0.7 / p.shape[1] ** 0.5)
    N = 5
    cond = torch.randn(1, 6, 32, generator=gen)
    emb_masks = torch.tensor([[0, 0, 1, 1, 1, 1]], dtype=torch.float32)
    cond = cond * emb_masks[:, :, None]
    torch.manual_seed(11)
    with Recorder() as rec, torch.no_grad():
        seq = ref.generate_video_diff.generate(model, cond, N, emb_masks=emb_masks, cfg_scale=1.0, temperature=0.9,
                                               cfg_iter=1.0)
    assert len(rec.draws) == N * 101
    noise, step_noise = [], []
    for j in range(N):
        a, b = split_draws(rec.draws[j * 101:(j + 1) * 101], 100)
        noise.append(a)
        step_noise.append(b)
    out = {("w." + k): v.numpy() for k, v in model.state_dict().items() if "kv_cache" not in k}
    out.update(cond=cond.numpy(), emb_masks=emb_masks.numpy(), noise=torch.stack(noise).numpy(),
               step_noise=torch.stack(step_noise).numpy(), temperature=np.float32(0.9), seq=seq.numpy())
    print("generate seq", tuple(seq.shape), "rms", float(seq.pow(2).mean().sqrt()))
    save("diff_generate.npz", **out)


if __name__ == "__main__":
    ref = _reference.load(diff=True)
    head_fixture(ref)
    generate_fixture(ref)
