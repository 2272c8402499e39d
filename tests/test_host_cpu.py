"""CPU: host-side logic of the drop-in modules (no kernels run here)."""
import numpy as np
import pytest
import torch

from vlg import _lib
from vlg.gpt import GPT_models, ModelArgs, Transformer, precompute_freqs_cis_2d, precompute_freqs_cis_3d_video
from vlg.vq_model import VectorQuantizer
from oracle import ar_decode as O


def test_same_seed_same_weights_as_reference(golden):
    """The facade's constructor consumes torch's RNG exactly like the reference's (gpt.py:262-316)."""
    f = golden("gptb_greedy.npz")
    torch.manual_seed(0)
    m = GPT_models["GPT-B"](block_size=256, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(m.output.weight, std=0.02)
    sd = m.state_dict()
    assert {"sum." + k for k in sd} == {k for k in f.files if k.startswith("sum.")}
    for k, v in sd.items():
        s = f["sum." + k]
        assert abs(float(v.double().sum()) - s[0]) < 1e-9 and abs(float(v.double().abs().sum()) - s[1]) < 1e-9, k


def test_state_dict_keys_t2i():
    m = Transformer(ModelArgs(dim=128, n_layer=1, n_head=2, vocab_size=64, block_size=16, multiple_of=64,
                              cls_token_num=120, caption_dim=32, model_type="t2i"))
    keys = set(m.state_dict())
    assert {"cls_embedding.cap_proj.fc1.weight", "cls_embedding.cap_proj.fc2.weight", "cls_embedding.uncond_embedding",
            "tok_embeddings.weight", "layers.0.attention.wqkv.weight", "layers.0.feed_forward.w3.weight", "norm.weight",
            "output.weight"} <= keys
    assert m.cls_embedding.uncond_embedding.shape == (120, 32)
    assert float(m.output.weight.abs().max()) == 0.0                  # zero-initialised head, gpt.py:307


def test_rope_tables_match_oracle():
    assert torch.equal(precompute_freqs_cis_2d(24, 64, 10000, 1), O.rope_table_2d(24, 64, 10000.0, 1))
    assert torch.equal(precompute_freqs_cis_3d_video(4, 2, 64, 10000, 120), O.rope_table_3d(4, 2, 64, 10000.0, 120))


def test_compute_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = Transformer(ModelArgs(dim=128, n_layer=1, n_head=2, vocab_size=64, block_size=16, multiple_of=64,
                              num_classes=4)).eval()
    from vlg.generate import generate, sample
    with pytest.raises(_lib.VlgError):
        generate(m, torch.tensor([1, 2]), 4, cfg_scale=2.0)
    with pytest.raises(_lib.VlgError):
        sample(torch.zeros(2, 1, 64))
    vq = VectorQuantizer(64, 8, 0.25, 0.0, True, True).eval()
    with pytest.raises(_lib.VlgError):
        vq(torch.zeros(1, 8, 2, 2))
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 1, dtype=torch.long), torch.zeros(1, dtype=torch.long), targets=torch.zeros(1, 1))


def test_drop_in_import_paths():
    from autoregressive.models.gpt import GPT_models as G
    from autoregressive.models.generate import generate, sample, top_k_top_p_filtering  # noqa: F401
    from tokenizer.tokenizer_image.vq_model import VectorQuantizer as V  # noqa: F401
    assert set(G) == {"GPT-B", "GPT-L", "GPT-XL", "GPT-XXL", "GPT-XXXL", "GPT-1B", "GPT-3B", "GPT-7B"}
    from autoregressive.models.gpt_video import GPT_models as GV
    from autoregressive.models.generate_video import generate as generate_video   # the module the fork forgot to ship
    assert set(GV) == set(G) and callable(generate_video)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) on a 1-decode-step sample: one JSON line on stdout
    with the contract's keys, nothing else."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-decode-steps", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("cfg-2") and d["metric"].startswith("generated visual tokens/sec")


def test_sample_script_import_lines_resolve_to_the_drop_ins():
    """The import lines of the reference's samplers that belong to the path (autoregressive/sample/sample_t2i.py:24-27,
    serve/sample_c2i.py:6; `language.t5` on :25 is the reference's own T5 wrapper and stays the reference's)."""
    from tokenizer.tokenizer_image.vq_model import VQ_models
    from autoregressive.models.gpt import GPT_models
    from autoregressive.models.generate import generate
    import vlg.vq_model, vlg.gpt, vlg.generate
    assert VQ_models is vlg.vq_model.VQ_models and set(VQ_models) == {"VQ-16", "VQ-8"}
    assert GPT_models is vlg.gpt.GPT_models and generate is vlg.generate.generate
    vq = VQ_models["VQ-8"](codebook_size=64, codebook_embed_dim=8)          # ctor kwargs of sample_t2i.py:41-43
    assert {"quantize.embedding.weight", "quantize.codebook_used", "quant_conv.weight", "post_quant_conv.bias",
            "encoder.conv_in.weight", "decoder.conv_out.bias", "decoder.conv_blocks.0.attn.2.proj_out.weight",
            "decoder.conv_blocks.1.upsample.conv.weight", "encoder.conv_blocks.0.downsample.conv.bias",
            "encoder.mid.1.q.weight", "decoder.conv_blocks.3.res.0.nin_shortcut.weight"} <= set(vq.state_dict())
    assert vq.quantize.codebook_used.shape == (65536,)
    for name in ("encode", "decode", "decode_code", "forward"):
        assert callable(getattr(vq, name))


def test_diffusion_sampler_import_lines_resolve_to_the_drop_ins():
    """The import lines of the fork's diffusion-head sampler that belong to the path (autoregressive/sample/
    sample_t2v_1f_diff.py:28-29; gpt_video_diff.py:27 imports DiffLoss from a bare `diffloss` module in its own directory)."""
    from autoregressive.models.generate_video_diff import generate
    from autoregressive.models.gpt_video_diff import GPT_models
    from autoregressive.models.diffloss import DiffLoss
    import vlg.generate_video_diff, vlg.gpt_video_diff, vlg.diffloss
    assert generate is vlg.generate_video_diff.generate and GPT_models is vlg.gpt_video_diff.GPT_models
    assert DiffLoss is vlg.diffloss.DiffLoss and set(GPT_models) >= {"GPT-B", "GPT-L", "GPT-XL", "GPT-1B"}
    # ctor kwargs of sample_t2v_1f_diff.py:86-97 on a small registry-shaped model
    from vlg.gpt_video_diff import Transformer, ModelArgs
    m = Transformer(ModelArgs(dim=128, n_layer=1, n_head=2, block_size=16, cls_token_num=120, caption_dim=32, vae_embed_dim=8,
                              num_frames=1, model_type="t2v", vocab_size=64, diffloss_w=128, multiple_of=64))
    assert m.model_type == "t2v" and m.diffloss.in_channels == 8 and m.seq_len == 16
    with pytest.raises(NotImplementedError):
        generate(m, torch.zeros(1, 120, 32), 4, cfg_scale=4.0)             # the reference's broken branch is refused, not invented


def test_vqmodel_same_seed_same_weights_and_cnn_parity(golden):
    """VQ_models['VQ-16'] consumes torch's RNG like the reference's constructor (same keys, same weights per seed), and
    its encoder / decoder CNNs (library convolutions) reproduce the reference module's outputs on CPU fp32."""
    f = golden("vqmodel.npz")
    from vlg.vq_model import VQ_models
    torch.manual_seed(0)
    m = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).eval()
    sd = m.state_dict()
    assert {"sum." + k for k in sd} == {k for k in f.files if k.startswith("sum.")}
    for k, v in sd.items():
        s = f["sum." + k]
        assert abs(float(v.double().sum()) - s[0]) < 1e-6 and abs(float(v.double().abs().sum()) - s[1]) < 1e-6, k
    with torch.no_grad():
        h = m.quant_conv(m.encoder(torch.from_numpy(f["x"])))
        img = m.decoder(torch.from_numpy(f["feat"]))
    assert np.abs(h.numpy() - f["h"]).max() < 1e-4 * max(1.0, np.abs(f["h"]).max())
    assert np.abs(img.numpy() - f["img"]).max() < 1e-4 * max(1.0, np.abs(f["img"]).max())
    if not torch.cuda.is_available():
        with pytest.raises(_lib.VlgError):
            m.decode_code(torch.from_numpy(f["codes"]), (2, 8, 8, 8))
