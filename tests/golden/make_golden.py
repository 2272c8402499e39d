"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so
these files -- outputs of the unmodified reference modules on seeded inputs -- are
what pins ``oracle/`` and, through it, the CUDA path.  Each .npz records the torch
version that produced it.  Small geometries only (a few MB in total); the full-size
cases are covered by regenerating weights from a seed (``gptb_*``) and by
size-independent properties in the GPU tests.
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


def sd_numpy(model):
    return {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()
            if "kv_cache" not in k}


def save(name, **arrays):
    arrays["torch_version"] = np.array(torch.__version__)
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def teacher_forced(ref, model, rows, tokens, emb_masks=None):
    """Raw fp32 logits [N,R,V] from the reference module stepped by hand."""
    T = 1 if model.model_type == "c2i" else rows.shape[1]
    N = tokens.shape[1] + 1
    model.setup_caches(max_batch_size=rows.shape[0], max_seq_length=T + N,
                       dtype=model.tok_embeddings.weight.dtype)
    if emb_masks is not None:                   # generate.py:156-165
        m = torch.cat([emb_masks, emb_masks])
        model.causal_mask[:, :, :T] = model.causal_mask[:, :, :T] * m.unsqueeze(1)
        eye = torch.eye(model.causal_mask.size(1), model.causal_mask.size(2))
        model.causal_mask[:] = model.causal_mask * (1 - eye) + eye
    out = []
    with torch.no_grad():
        logits, _ = model(None, rows, torch.arange(T))
        out.append(logits[:, -1])
        for i in range(N - 1):
            x = torch.cat([tokens[:, i:i + 1], tokens[:, i:i + 1]])
            with torch.backends.cuda.sdp_kernel(enable_flash=False, enable_mem_efficient=False, enable_math=True):
                logits, _ = model(x, None, torch.tensor([T + i]))
            out.append(logits[:, -1])
    return torch.stack(out)


def c2i_tiny(ref):
    torch.manual_seed(7)
    args = ref.gpt.ModelArgs(dim=128, n_layer=2, n_head=2, vocab_size=512, block_size=16, multiple_of=64,
                             cls_token_num=1, num_classes=10, model_type="c2i")
    model = ref.gpt.Transformer(args).eval()
    torch.nn.init.normal_(model.output.weight, std=0.05)
    torch.nn.init.normal_(model.tok_embeddings.weight, std=0.5)      # "rich" init: greedy does not collapse
    torch.nn.init.normal_(model.cls_embedding.embedding_table.weight, std=0.5)
    for n, p in model.named_parameters():
        if n.endswith("norm.weight"):
            torch.nn.init.normal_(p, mean=1.0, std=0.1)
    cond = torch.tensor([1, 7, 3])
    greedy = ref.generate.generate(model, cond, 16, cfg_scale=4.0, temperature=1.0, top_k=1, top_p=1.0,
                                   sample_logits=False)
    greedy_nocfg = ref.generate.generate(model, cond, 16, cfg_scale=1.0, temperature=1.0, top_k=0, top_p=1.0,
                                         sample_logits=False)
    greedy_interval = ref.generate.generate(model, cond, 16, cfg_scale=3.0, cfg_interval=5, temperature=1.0,
                                            top_k=0, top_p=1.0, sample_logits=False)
    g = torch.Generator().manual_seed(11)
    tokens = torch.randint(0, 512, (3, 15), generator=g)
    rows = torch.cat([cond, torch.ones_like(cond) * 10])
    logits = teacher_forced(ref, model, rows, tokens)
    save("c2i_tiny.npz", cond=cond.numpy(), greedy=greedy.numpy(), greedy_nocfg=greedy_nocfg.numpy(),
         greedy_interval=greedy_interval.numpy(), tokens=tokens.numpy(), logits=logits.numpy(),
         **{"w." + k: v for k, v in sd_numpy(model).items()})


def t2i_tiny(ref):
    torch.manual_seed(8)
    args = ref.gpt.ModelArgs(dim=128, n_layer=2, n_head=2, vocab_size=256, block_size=16, multiple_of=64,
                             cls_token_num=120, caption_dim=32, model_type="t2i")   # uncond_embedding is [120, .] (gpt.py:93)
    model = ref.gpt.Transformer(args).eval()
    torch.nn.init.normal_(model.output.weight, std=0.05)
    torch.nn.init.normal_(model.tok_embeddings.weight, std=0.5)
    torch.nn.init.normal_(model.cls_embedding.cap_proj.fc2.weight, std=0.3)
    g = torch.Generator().manual_seed(12)
    cond = torch.randn(3, 120, 32, generator=g)
    valid = [120, 9, 57]                                              # left-padded: ones at the END
    emb_masks = torch.zeros(3, 120)
    for b, n in enumerate(valid):
        emb_masks[b, 120 - n:] = 1
    cond = cond * emb_masks[:, :, None]
    greedy = ref.generate.generate(model, cond, 16, emb_masks, cfg_scale=4.0, temperature=1.0, top_k=1,
                                   top_p=1.0, sample_logits=False)
    tokens = torch.randint(0, 256, (3, 15), generator=g)
    rows = torch.cat([cond, torch.zeros_like(cond) + model.cls_embedding.uncond_embedding])
    logits = teacher_forced(ref, model, rows, tokens, emb_masks)
    save("t2i_tiny.npz", cond=cond.numpy(), emb_masks=emb_masks.numpy(), greedy=greedy.numpy(),
         tokens=tokens.numpy(), logits=logits.numpy(), **{"w." + k: v for k, v in sd_numpy(model).items()})


def t2v_tiny(refv):
    gv = refv.gpt_video
    torch.manual_seed(9)
    args = gv.ModelArgs(dim=128, n_layer=2, n_head=2, vocab_size=64, block_size=16, multiple_of=64,
                        cls_token_num=120, caption_dim=64, model_type="t2v", vae_embed_dim=8, num_frames=5,
                        t_downsample_size=4, shuffle_video_tokens=False)
    model = gv.Transformer(args).eval()
    for n, p in model.named_parameters():
        if "vae_latent_adapter" in n or "cap_proj.fc2" in n:
            torch.nn.init.normal_(p, std=0.3)
    g = torch.Generator().manual_seed(13)
    cond = torch.randn(2, 120, 64, generator=g)
    emb_masks = torch.zeros(2, 120)
    emb_masks[0, 20:] = 1
    emb_masks[1, 113:] = 1
    cond = cond * emb_masks[:, :, None]
    rows = torch.cat([cond, torch.zeros_like(cond) + model.cls_embedding.uncond_embedding])
    N, T = 32, 120                                                     # 2 frames x 16 latents
    model.setup_caches(max_batch_size=4, max_seq_length=T + N, dtype=model.tok_embeddings.weight.dtype)
    m = torch.cat([emb_masks, emb_masks])
    model.causal_mask[:, :, :T] = model.causal_mask[:, :, :T] * m.unsqueeze(1)
    eye = torch.eye(model.causal_mask.size(1), model.causal_mask.size(2))
    model.causal_mask[:] = model.causal_mask * (1 - eye) + eye
    latents = torch.randn(2, N - 1, 8, generator=g)                   # teacher-forced inputs
    out = []
    with torch.no_grad():
        h, _ = model(idx=None, cond_embed=rows, input_pos=torch.arange(T))
        out.append(h[:, -1])
        for i in range(N - 1):
            x = torch.cat([latents[:, i:i + 1], latents[:, i:i + 1]])
            h, _ = model(idx=None, cond_embed=None, video_latent=x, input_pos=torch.tensor([T + i]))
            out.append(h[:, -1])
    save("t2v_tiny.npz", cond=cond.numpy(), emb_masks=emb_masks.numpy(), latents=latents.numpy(),
         out=torch.stack(out).numpy(), rope=model.freqs_cis.numpy(),
         **{"w." + k: v for k, v in sd_numpy(model).items()})


def t2v_generate(refv):
    """The t2v generation protocol of vlg.generate_video / oracle.generate_video driven through the REFERENCE model's own
    forward (gpt_video.py:403-431) on the t2v_tiny weights: prefill, CFG mix of the output halves (generate.py:95-97
    formula), the mixed latent is the next input.  models/generate_video.py itself is missing from the fork."""
    gv = refv.gpt_video
    f = np.load(os.path.join(HERE, "t2v_tiny.npz"))
    args = gv.ModelArgs(dim=128, n_layer=2, n_head=2, vocab_size=64, block_size=16, multiple_of=64,
                        cls_token_num=120, caption_dim=64, model_type="t2v", vae_embed_dim=8, num_frames=5,
                        t_downsample_size=4, shuffle_video_tokens=False)
    model = gv.Transformer(args).eval()
    model.load_state_dict({k[2:]: torch.from_numpy(f[k]) for k in f.files if k.startswith("w.")}, strict=True)
    cond, emb_masks = torch.from_numpy(f["cond"]), torch.from_numpy(f["emb_masks"])
    rows = torch.cat([cond, torch.zeros_like(cond) + model.cls_embedding.uncond_embedding])
    N, T, B = 24, 120, cond.shape[0]
    out = {}
    for name, scale, interval in (("gen", 4.0, -1), ("gen_interval", 4.0, 3)):
        model.setup_caches(max_batch_size=2 * B, max_seq_length=T + N, dtype=model.tok_embeddings.weight.dtype)
        m = torch.cat([emb_masks, emb_masks])
        model.causal_mask[:, :, :T] = model.causal_mask[:, :, :T] * m.unsqueeze(1)
        eye = torch.eye(model.causal_mask.size(1), model.causal_mask.size(2))
        model.causal_mask[:] = model.causal_mask * (1 - eye) + eye
        with torch.no_grad():
            h, _ = model(idx=None, cond_embed=rows, input_pos=torch.arange(T))
            c, u = torch.split(h[:, -1], B, dim=0)
            lat = u + (c - u) * scale
            seq = [lat]
            guided = True
            for i in range(N - 1):
                if interval > -1 and i > interval:
                    guided = False
                x = torch.cat([lat, lat]).unsqueeze(1)
                h, _ = model(idx=None, cond_embed=None, video_latent=x, input_pos=torch.tensor([T + i]))
                c, u = torch.split(h[:, -1], B, dim=0)
                lat = u + (c - u) * scale if guided else c
                seq.append(lat)
        out[name] = torch.stack(seq, dim=1).numpy()
    save("t2v_generate.npz", **out)


def sampling(ref):
    g = torch.Generator().manual_seed(21)
    logits = torch.randn(5, 1, 1024, generator=g) * 3
    logits[1, 0, 100:110] = logits[1, 0, 100]                        # exact ties around a threshold
    out = {"logits": logits.numpy()}
    cases = {"k50": dict(top_k=50), "k1": dict(top_k=1), "p80": dict(top_p=0.8),
             "k100_p90_t07": dict(top_k=100, top_p=0.9, temperature=0.7), "plain": dict(),
             "k2000": dict(top_k=2000), "t0": dict(temperature=0.0, top_k=5)}
    for name, kw in cases.items():
        idx, probs = ref.generate.sample(logits.clone(), sample_logits=False, **kw)
        out[name + ".idx"] = idx.numpy()
        out[name + ".probs"] = probs.numpy()
    save("sampling.npz", **out)


def vq(ref):
    torch.manual_seed(0)
    q = ref.vq_model.VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval()
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(4, 8, 16, 16, generator=g)
    with torch.no_grad():
        z_q, _, (_, _, idx) = q(z)
        entry = q.get_codebook_entry(idx, (4, 8, 16, 16), channel_first=True)
    # a second, un-normalised codebook exercising l2_norm=False and a small table
    torch.manual_seed(1)
    q2 = ref.vq_model.VectorQuantizer(1000, 8, 0.25, 0.0, False, False).eval()
    q2.embedding.weight.data.normal_()
    z2 = torch.randn(2, 8, 5, 7, generator=g)
    with torch.no_grad():
        z_q2, _, (_, _, idx2) = q2(z2)
    save("vq.npz", codebook=q.embedding.weight.detach().numpy(), z=z.numpy(), z_q=z_q.numpy(), idx=idx.numpy(),
         entry=entry.numpy(), codebook2=q2.embedding.weight.detach().numpy(), z2=z2.numpy(),
         z_q2=z_q2.numpy(), idx2=idx2.numpy())


def gptb(ref):
    """BASELINE cfg-1 at its stated size: GPT-B c2i 256 px (16x16 = 256 tokens), batch 4, CFG 4.0, fp32, the reference on
    CPU -- greedy (top-k 1) so that the token sequences are the bit-exact contract.  Weights are NOT stored; they are
    re-created from the seed by the product's own constructor, whose parameter-creation order must therefore match the
    reference's (checksums below).  `gap` is the oracle's top-1 / top-2 distance of the CFG-mixed fp32 logits at every
    step of that sequence: the diagnostic to print when a kernel diverges (hard part 1)."""
    from oracle import ar_decode as O
    torch.manual_seed(0)
    model = ref.gpt.GPT_models["GPT-B"](block_size=256, cls_token_num=1, model_type="c2i").eval()
    torch.nn.init.normal_(model.output.weight, std=0.02)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items() if "kv_cache" not in k}
    sums = {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in sd.items()}
    cond = torch.randint(0, 1000, (4,), generator=torch.Generator().manual_seed(1234))
    greedy = ref.generate.generate(model, cond, 256, cfg_scale=4.0, temperature=1.0, top_k=1, top_p=1.0,
                                   sample_logits=False)
    geo = O.Geometry(block_size=256, **O.NAMED["GPT-B"])
    logits = O.teacher_forced_logits(geo, sd, cond, greedy[:, :-1].long())          # [N, 2B, V] raw
    mixed = torch.stack([O.cfg_mix(l.unsqueeze(1), 4.0)[:, 0] for l in logits])      # [N, B, V]
    top2 = mixed.topk(2, dim=-1)
    assert torch.equal(top2.indices[..., 0].t().int(), greedy), "oracle and reference disagree on the greedy sequence"
    gap = (top2.values[..., 0] - top2.values[..., 1]).t().contiguous()               # [B, N]
    save("gptb_greedy.npz", cond=cond.numpy(), greedy=greedy.numpy(), gap=gap.numpy(),
         **{"sum." + k: v for k, v in sums.items()})


def vqmodel(ref):
    """VQ-16 tokenizer (vq_model.py:28-62): encode() of a seeded image and decode_code() of seeded codes through the
    reference module on CPU fp32.  Weights are re-created from the seed (72 M parameters are not stored): checksums pin
    the construction order."""
    torch.manual_seed(0)
    m = ref.vq_model.VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).eval()
    sums = {k: np.array([float(v.double().sum()), float(v.double().abs().sum())]) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    x = torch.randn(2, 3, 64, 64, generator=g)
    with torch.no_grad():
        h = m.quant_conv(m.encoder(x))
        quant, _, (_, _, idx) = m.encode(x)
        codes = torch.randint(0, 16384, (2 * 8 * 8,), generator=g)
        img = m.decode_code(codes, (2, 8, 8, 8))
        feat = m.post_quant_conv(m.quantize.get_codebook_entry(codes, (2, 8, 8, 8)))
    save("vqmodel.npz", x=x.numpy(), h=h.numpy(), quant=quant.numpy(), idx=idx.numpy(), codes=codes.numpy(),
         img=img.numpy(), feat=feat.numpy(), **{"sum." + k: v for k, v in sums.items()})


if __name__ == "__main__":
    ref = _reference.load()
    which = sys.argv[1:] or ["c2i", "t2i", "t2v", "sampling", "vq", "gptb", "vqmodel"]
    with torch.no_grad():
        if "c2i" in which:
            c2i_tiny(ref)
        if "t2i" in which:
            t2i_tiny(ref)
        if "sampling" in which:
            sampling(ref)
        if "vq" in which:
            vq(ref)
        if "gptb" in which:
            gptb(ref)
        if "vqmodel" in which:
            vqmodel(ref)
        if "t2v" in which:
            t2v_tiny(_reference.load(video=True))
        if "t2vgen" in which:
            t2v_generate(_reference.load(video=True))
