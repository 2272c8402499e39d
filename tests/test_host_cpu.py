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
