"""GPU: the diffusion-head sampler (vlg_diff_sample behind vlg.diffloss.DiffLoss.sample) and the generate_video_diff loop
against reference-generated fixtures and the oracle running in fp32 on the same device.

Tolerance: fp32 end to end on both sides; what differs is the summation order of the 1024-long dot products and of the
LayerNorm statistics (~1e-7 relative per layer), carried through 100 dependent steps whose measured amplification is of
order one (oracle in fp64 against fp32: 5e-7 of the output's absmax).  The gate is 2e-4 of the output's absmax.
"""
import numpy as np
import pytest
import torch

from oracle import ar_decode as A, diffusion as D
from vlg import _lib as L
from vlg.diffloss import DiffLoss
from vlg import gpt_video_diff as VD
from vlg.generate_video_diff import generate

pytestmark = pytest.mark.gpu
DEV = "cuda"
REL = 2e-4


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def fixture_head(f):
    d = DiffLoss(target_channels=8, z_channels=96, depth=3, width=128, num_sampling_steps="100").eval()
    d.load_state_dict({k[2:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("w.")}, strict=True)
    return d.to(DEV)


@pytest.mark.parametrize("tag", ["plain", "temp", "cfg"])
def test_fixture_chain_matches_the_reference(golden, tag):
    """The reference's own output on the reference's own noise draws (tests/golden/make_golden_diff.py)."""
    f = golden("diff_head.npz")
    d = fixture_head(f)
    out = d.sample(torch.from_numpy(f[tag + ".z"]).to(DEV), float(f[tag + ".temperature"]), float(f[tag + ".cfg"]),
                   noise=torch.from_numpy(f[tag + ".noise"]), step_noise=torch.from_numpy(f[tag + ".step_noise"]))
    d.engine(1).check()
    ref = torch.from_numpy(f[tag + ".sample"])
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_err(out.cpu(), ref) < REL, rel_err(out.cpu(), ref)


def trained_like(module, gen):
    for m in module.modules():
        if isinstance(m, torch.nn.Linear):
            m.weight.data = torch.randn(m.weight.shape, generator=gen) * (1.0 / m.weight.shape[1] ** 0.5)
            m.bias.data = torch.randn(m.bias.shape, generator=gen) * 0.05
        elif isinstance(m, torch.nn.LayerNorm) and m.elementwise_affine:
            m.weight.data = 1.0 + 0.1 * torch.randn(m.weight.shape, generator=gen)
            m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=gen)


@pytest.mark.parametrize("width,depth,zdim,rows,cfg", [
    (1024, 3, 2048, 1, 1.0),      # the shipped configuration: GPT-1B trunk, batch 1, no guidance
    (1024, 3, 2048, 2, 2.5),      # guidance inside the head (cond row, uncond row)
    (1024, 3, 1280, 5, 1.0),      # a ragged row count (two register passes, the second one short)
    (1024, 3, 1280, 16, 3.0),     # the row limit of one launch
    (512, 6, 768, 3, 1.0),        # deeper and narrower: four columns per CTA
    (256, 2, 64, 4, 1.0),
])
def test_full_size_chain_matches_the_oracle(width, depth, zdim, rows, cfg):
    gen = torch.Generator().manual_seed(width + depth + rows)
    d = DiffLoss(target_channels=8, z_channels=zdim, depth=depth, width=width, num_sampling_steps="100").eval()
    trained_like(d, gen)
    z = torch.randn(rows, zdim, generator=gen)
    half = rows // 2
    noise = torch.randn(rows, 8, generator=gen)
    if cfg != 1.0:
        noise = torch.cat([noise[:half], noise[:half]])
    step_noise = torch.randn(100, rows, 8, generator=gen)
    d = d.to(DEV)
    out = d.sample(z.to(DEV), 0.9, cfg, noise=noise, step_noise=step_noise)
    d.engine(1).check()
    w = {k: v.to(DEV) for k, v in d.state_dict().items()}
    ref = D.p_sample_loop(w, z.to(DEV), noise.to(DEV), step_noise.to(DEV), 0.9, cfg)
    assert torch.isfinite(out).all()
    assert rel_err(out, ref) < REL, rel_err(out, ref)


def test_other_step_counts_and_more_rows_than_one_launch():
    gen = torch.Generator().manual_seed(77)
    d = DiffLoss(target_channels=4, z_channels=256, depth=2, width=256, num_sampling_steps="50").eval()
    trained_like(d, gen)
    z = torch.randn(21, 256, generator=gen)                      # > 16 rows: chunked by the host wrapper
    noise = torch.randn(21, 4, generator=gen)
    step_noise = torch.randn(50, 21, 4, generator=gen)
    d = d.to(DEV)
    out = d.sample(z.to(DEV), 1.0, 1.0, noise=noise, step_noise=step_noise)
    w = {k: v.to(DEV) for k, v in d.state_dict().items()}
    ref = D.p_sample_loop(w, z.to(DEV), noise.to(DEV), step_noise.to(DEV), 1.0, 1.0, num_sampling_steps="50")
    assert rel_err(out, ref) < REL


def test_in_kernel_noise_is_seeded_and_gaussian():
    """No supplied draws: Philox in the kernel.  With a zeroed output layer the chain is linear in the draws, so the result is
    a zero-mean Gaussian whose standard deviation the oracle gives by Monte Carlo on torch's generator."""
    gen = torch.Generator().manual_seed(5)
    d = DiffLoss(target_channels=8, z_channels=64, depth=1, width=128, num_sampling_steps="100").eval()
    trained_like(d, gen)
    torch.nn.init.zeros_(d.net.final_layer.linear.weight)
    torch.nn.init.zeros_(d.net.final_layer.linear.bias)
    d = d.to(DEV)
    z = torch.zeros(16, 64, device=DEV)
    a = d.sample(z, 1.0, 1.0, seed=123)
    d._calls -= 1
    b = d.sample(z, 1.0, 1.0, seed=123)                          # same seed, same call offset: same draws
    assert torch.equal(a, b)
    c = d.sample(z, 1.0, 1.0, seed=124)
    assert not torch.equal(a, c)
    ours = torch.cat([d.sample(z, 1.0, 1.0, seed=1000 + i) for i in range(64)]).flatten()           # 8192 draws
    w = {k: v.to(DEV) for k, v in d.state_dict().items()}
    zz = torch.zeros(1024, 64, device=DEV)
    ref = D.p_sample_loop(w, zz, torch.randn(1024, 8, device=DEV), torch.randn(100, 1024, 8, device=DEV), 1.0, 1.0).flatten()
    assert abs(float(ours.mean())) < 4 * float(ref.std()) / 8192 ** 0.5
    assert abs(float(ours.std()) / float(ref.std()) - 1.0) < 0.05
    # rows are independent draws, not copies of one another
    assert float(torch.corrcoef(torch.stack([a[0], a[1]]))[0, 1].abs()) < 0.95


def test_geometry_limits_fail_loudly():
    d = DiffLoss(target_channels=8, z_channels=64, depth=3, width=384, num_sampling_steps="100").to(DEV)
    with pytest.raises(L.VlgError, match="unsupported"):
        d.sample(torch.zeros(1, 64, device=DEV))
    d = DiffLoss(target_channels=8, z_channels=64, depth=8, width=1024, num_sampling_steps="100").to(DEV)
    with pytest.raises(L.VlgError, match="shared memory"):
        d.sample(torch.zeros(1, 64, device=DEV))
    d = DiffLoss(target_channels=8, z_channels=64, depth=1, width=128, num_sampling_steps="100").to(DEV)
    with pytest.raises(L.VlgError, match="even row count"):
        d.sample(torch.zeros(3, 64, device=DEV), 1.0, 2.0)
    with pytest.raises(L.VlgError, match="float32"):
        DiffLoss(8, 64, 1, 128, "100").to(DEV, torch.bfloat16).sample(torch.zeros(1, 64, device=DEV))


def test_generate_video_diff_loop_matches_the_reference(golden):
    """generate_video_diff.generate of the reference (cfg_scale 1, B = 1, masked caption prefix) on the tiny trunk + head:
    trunk positions through vlg_forward_pos in fp32, every token's latent through vlg_diff_sample, the reference's draws."""
    g = golden("diff_generate.npz")
    m = VD.Transformer(VD.ModelArgs(dim=128, n_layer=2, n_head=2, block_size=16, cls_token_num=6, caption_dim=32,
                                    vae_embed_dim=8, num_frames=5, t_downsample_size=4, vocab_size=64, diffloss_w=128,
                                    diffloss_d=3, multiple_of=64)).eval()
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}, strict=True)
    m = m.to(DEV)
    ref = torch.from_numpy(g["seq"])
    seq = generate(m, torch.from_numpy(g["cond"]).to(DEV), ref.shape[1], emb_masks=torch.from_numpy(g["emb_masks"]).to(DEV),
                   cfg_scale=1.0, temperature=float(g["temperature"]), cfg_iter=1.0,
                   noise=torch.from_numpy(g["noise"]), step_noise=torch.from_numpy(g["step_noise"]))
    assert seq.shape == ref.shape
    # each token's latent feeds the next trunk position, so the head's 2e-4 compounds over the 5 tokens
    assert rel_err(seq.cpu(), ref) < 5 * REL, rel_err(seq.cpu(), ref)
    with pytest.raises(NotImplementedError):
        generate(m, torch.from_numpy(g["cond"]).to(DEV), 2, cfg_scale=4.0)


def test_forward_prefix_equals_stepped_prefill(golden):
    """vlg_forward_prefix (all caption positions in one pass) against stepping them one by one through vlg_forward_pos:
    same hidden state at the last position, same K/V in the caches (checked through the next position's output)."""
    g = golden("diff_generate.npz")
    m = VD.Transformer(VD.ModelArgs(dim=128, n_layer=2, n_head=2, block_size=16, cls_token_num=6, caption_dim=32,
                                    vae_embed_dim=8, num_frames=5, t_downsample_size=4, vocab_size=64, diffloss_w=128,
                                    diffloss_d=3, multiple_of=64)).eval()
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}, strict=True)
    m = m.to(DEV)
    cond = torch.from_numpy(g["cond"]).to(DEV)
    masks = torch.from_numpy(g["emb_masks"]).to(DEV)
    lat = torch.randn(1, 1, 8, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    outs = []
    for batched in (False, True):
        m.setup_caches(1, 6 + 4, None)
        m.set_prefix_mask(masks)
        if batched:
            eng = m._engine
            h_last = eng.forward_prefix(eng.mlp(0, cond, 128)[:, :6])[1]
        else:
            h_last = m(cond_embed=cond, input_pos=torch.arange(6))[0][:, -1]
        h_next = m(video_latent=lat, input_pos=torch.tensor([6]))[0][:, -1]
        outs.append((h_last.clone(), h_next.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), float((a - b).abs().max())
