"""Drop-in for the reference's ``autoregressive/models/generate_video_diff.py``: the t2v loop whose every token is drawn
by the diffusion head (the fork's actual t2v sampler, ``sample_t2v_1f_diff.py:162-170``).

    token_embedding_sample = generate(gpt_model, c_indices, vae_t * latent_size ** 2, c_emb_masks,
                                      cfg_scale=args.cfg_scale, temperature=args.temperature, cfg_iter=args.cfg_iter)

Per token: one trunk position (``vlg_forward_pos``; the prefix in one go at prefill) -> final-normed hidden state ->
``DiffLoss.sample`` (``vlg_diff_sample``: the 100-step chain in one launch) -> latent, which is the step's output and,
through ``vae_latent_adapter``, the next input (generate_video_diff.py:77-95, 108-137, 175-212).

Only ``cfg_scale == 1`` exists in the reference: with ``cfg_scale > 1`` its ``decode_one_token`` computes the guided hidden
state and then returns a variable it never assigned (generate_video_diff.py:112-124, 137), and the shipped script passes
``--cfg-scale 1``.  So ``cfg_scale > 1`` raises here instead of inventing behaviour; guidance INSIDE the head
(``cfg_iter``, forwarded to ``DiffLoss.sample(cfg=...)``) is supported as the reference defines it.  The reference
folds the batch into the sequence axis (``.view(1, -1, C)``, :93), i.e. it runs with B == 1; rows stay rows here.
"""
import torch

from . import _lib as L


@torch.no_grad()
def generate(model, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, temperature=1.0, cfg_iter=1.0,
             noise=None, step_noise=None, **sampling_kwargs):
    """-> latents [B, max_new_tokens, vae_embed_dim] in ``cond.dtype`` (generate_video_diff.py:175-212).

    ``noise`` [max_new_tokens, B, C] / ``step_noise`` [max_new_tokens, steps, B, C] replace the in-kernel draws (parity tests)."""
    if model.model_type != 't2v':
        raise Exception("please check model type")
    if cfg_scale > 1.0:
        raise NotImplementedError("generate_video_diff with cfg_scale > 1 has no defined behaviour in the reference "
                                  "(decode_one_token never assigns its result, generate_video_diff.py:112-137); use cfg_iter")
    T = cond.shape[1]
    assert T == model.cls_token_num
    dtype = model.tok_embeddings.weight.dtype
    cond_in = cond.to(dtype).contiguous()
    L.require_cuda(cond_in, "generate")
    B = cond.shape[0]
    if emb_masks is not None:
        assert emb_masks.shape[0] == B and emb_masks.shape[-1] == T
    model.setup_caches(max_batch_size=B, max_seq_length=T + max_new_tokens, dtype=dtype)
    if emb_masks is not None:
        model.set_prefix_mask(emb_masks)
    head = model.diffloss
    seq = torch.empty(B, max_new_tokens, model.config.vae_embed_dim, dtype=cond.dtype, device=cond.device)

    def draw(hidden, j):
        return head.sample(hidden.float(), temperature, cfg_iter, noise=None if noise is None else noise[j],
                           step_noise=None if step_noise is None else step_noise[j])

    # prefill: the whole caption prefix in one pass (only its last position's hidden state is used, generate_video_diff.py:80)
    eng = model._engine
    emb = eng.mlp(0, cond_in, model.config.dim)[:, :T]
    lat = draw(eng.forward_prefix(emb)[1], 0)
    seq[:, 0] = lat
    for j in range(1, max_new_tokens):
        hidden, _ = model(video_latent=lat.to(dtype).unsqueeze(1), cond_embed=None, input_pos=torch.tensor([T + j - 1]))
        lat = draw(hidden[:, -1], j)
        seq[:, j] = lat
    return seq
