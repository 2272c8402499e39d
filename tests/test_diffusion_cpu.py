"""CPU: the diffusion-head oracle against the reference-generated fixtures (and the live reference when it is present),
and the host logic of the drop-in modules (schedule tables, module tree, seeded weights).  No kernels run here."""
import numpy as np
import pytest
import torch

from oracle import _reference, ar_decode as A, diffusion as D
from vlg.diffloss import DiffLoss, sampling_schedule
from vlg import gpt_video_diff as VD

TABLES = ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
          "posterior_log_variance_clipped", "log_betas")
TINY_DIFF = A.Geometry(dim=128, n_layer=2, n_head=2, vocab_size=64, block_size=16, multiple_of=64, cls_token_num=6,
                       caption_dim=32, model_type="t2v", vae_embed_dim=8, num_frames=5, t_downsample_size=4)


def head_weights(f):
    return {k[2:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("w.")}


def test_oracle_schedule_equals_the_reference_tables(golden):
    f = golden("diff_head.npz")
    t = D.sampling_tables("100")
    assert np.array_equal(t["timestep_map"], f["tab.timestep_map"])
    for name in TABLES:
        assert np.array_equal(t[name], f["tab." + name]), name            # float64, bit for bit


def test_host_schedule_equals_the_reference_tables(golden):
    """vlg.diffloss.sampling_schedule is what the product hands to vlg_diff_create (as fp32, like _extract_into_tensor)."""
    f = golden("diff_head.npz")
    tmap, tab = sampling_schedule("100")
    assert np.array_equal(tmap, f["tab.timestep_map"])
    for i, name in enumerate(TABLES):
        assert np.array_equal(tab[i].astype(np.float32), f["tab." + name].astype(np.float32)), name
    for spec in ("50", "250", "10,20,30", "1000"):
        tmap, tab = sampling_schedule(spec)
        ref = D.sampling_tables(spec)
        assert np.array_equal(tmap, ref["timestep_map"])
        for i, name in enumerate(TABLES):
            assert np.array_equal(tab[i].astype(np.float32), ref[name].astype(np.float32)), (spec, name)


def test_oracle_net_forward_is_the_reference(golden):
    f = golden("diff_head.npz")
    w = head_weights(f)
    x, t, c = (torch.from_numpy(f["net." + k]) for k in ("x", "t", "c"))
    assert torch.equal(D.net_forward(w, x, t, c), torch.from_numpy(f["net.out"]))
    assert torch.equal(D.net_forward_with_cfg(w, x, t, c, 3.0), torch.from_numpy(f["net.out_cfg"]))


@pytest.mark.parametrize("tag", ["plain", "temp", "cfg"])
def test_oracle_sampling_loop_is_the_reference(golden, tag):
    """100-step chain on the reference's own noise draws: bit-exact on CPU (same torch ops in the same order)."""
    f = golden("diff_head.npz")
    w = head_weights(f)
    out = D.p_sample_loop(w, torch.from_numpy(f[tag + ".z"]), torch.from_numpy(f[tag + ".noise"]),
                          torch.from_numpy(f[tag + ".step_noise"]), float(f[tag + ".temperature"]), float(f[tag + ".cfg"]))
    ref = torch.from_numpy(f[tag + ".sample"])
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-6


def test_oracle_generate_loop_is_the_reference(golden):
    """generate_video_diff.generate (cfg_scale 1, B = 1) on the tiny trunk + head fixture."""
    g = golden("diff_generate.npz")
    w = head_weights(g)
    seq = D.generate_video_diff(TINY_DIFF, w, torch.from_numpy(g["cond"]), g["seq"].shape[1], torch.from_numpy(g["emb_masks"]),
                                torch.from_numpy(g["noise"]), torch.from_numpy(g["step_noise"]), float(g["temperature"]))
    ref = torch.from_numpy(g["seq"])
    assert seq.shape == ref.shape
    assert float((seq - ref).abs().max() / ref.abs().max()) < 2e-6


def test_module_tree_matches_the_fixture_keys(golden):
    f = golden("diff_head.npz")
    d = DiffLoss(target_channels=8, z_channels=96, depth=3, width=128, num_sampling_steps="100")
    sd = d.state_dict()
    assert set(sd) == {k[2:] for k in f.files if k.startswith("w.")}
    for k, v in sd.items():
        assert tuple(v.shape) == f["w." + k].shape, k
    d.load_state_dict({k[2:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("w.")}, strict=True)
    g = golden("diff_generate.npz")
    m = VD.Transformer(VD.ModelArgs(dim=128, n_layer=2, n_head=2, block_size=16, cls_token_num=6, caption_dim=32,
                                    vae_embed_dim=8, num_frames=5, t_downsample_size=4, vocab_size=64, diffloss_w=128,
                                    diffloss_d=3, multiple_of=64))
    assert set(m.state_dict()) == {k[2:] for k in g.files if k.startswith("w.")}
    m.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w.")}, strict=True)


def test_no_cpu_fallback():
    d = DiffLoss(8, 32, 1, 128, "100")
    with pytest.raises(Exception, match="no CPU fallback|CPU"):
        d.sample(torch.zeros(1, 32))
    with pytest.raises(NotImplementedError):
        d(torch.zeros(1, 8), torch.zeros(1, 32))


@pytest.mark.skipif(not _reference.available(), reason="reference tree not present (authoring container only)")
def test_same_seed_same_weights_as_the_live_reference():
    ref = _reference.load(diff=True)
    torch.manual_seed(5)
    a = ref.diffloss.DiffLoss(target_channels=8, z_channels=64, depth=2, width=128, num_sampling_steps="100")
    torch.manual_seed(5)
    b = DiffLoss(target_channels=8, z_channels=64, depth=2, width=128, num_sampling_steps="100")
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    kw = dict(dim=128, n_layer=1, n_head=2, block_size=16, cls_token_num=6, caption_dim=32, vae_embed_dim=8, num_frames=5,
              vocab_size=64, diffloss_w=128, multiple_of=64)
    torch.manual_seed(6)
    ma = ref.gpt_video_diff.Transformer(ref.gpt_video_diff.ModelArgs(**kw))
    torch.manual_seed(6)
    mb = VD.Transformer(VD.ModelArgs(**kw))
    sa, sb = ma.state_dict(), mb.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


@pytest.mark.skipif(not _reference.available(), reason="reference tree not present (authoring container only)")
def test_oracle_against_the_live_reference_on_a_fresh_seed():
    """Not a fixture replay: new weights, new draws, the reference run here and now (patched only for .cuda() / randn capture)."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("mgd", os.path.join(os.path.dirname(__file__), "golden", "make_golden_diff.py"))
    mgd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgd)
    ref = _reference.load(diff=True)
    gen = torch.Generator().manual_seed(4242)
    d = ref.diffloss.DiffLoss(target_channels=4, z_channels=48, depth=2, width=128, num_sampling_steps="50").eval()
    mgd.trained_like(d, gen)
    z = torch.randn(2, 48, generator=gen)
    torch.manual_seed(99)
    with mgd.Recorder() as rec, torch.no_grad():
        sample = d.sample(z, 0.8, 1.0)
    noise, step_noise = mgd.split_draws(rec.draws, 50)
    out = D.p_sample_loop(dict(d.state_dict()), z, noise, step_noise, 0.8, 1.0, num_sampling_steps="50")
    assert float((out - sample).abs().max() / sample.abs().max()) < 1e-6
