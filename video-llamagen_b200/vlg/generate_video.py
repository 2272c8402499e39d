"""Drop-in for the module the reference's t2v samplers import as ``autoregressive.models.generate_video``
(``sample_t2v_1f.py:30``) -- the file itself is missing from the fork.  The call site fixes the
contract (``sample_t2v_1f.py:199-210``)::

    token_embedding_sample = generate(gpt_model, c_indices, vae_t * latent_size ** 2, c_emb_masks,
                                      cfg_scale=..., temperature=..., top_k=..., top_p=..., sample_logits=True)
    token_embedding_sample.view(-1, vae_t, latent_size, latent_size, token_embedding_sample.shape[-1])

i.e. ``generate`` returns the continuous latents ``[B, N, vae_embed_dim]`` the ``gpt_video`` trunk emits
(``gpt_video.py:431``: no logits, no sampler).  The loop is generate.py:127-180 with the token hand-off replaced by
the latent hand-off: prefill over the caption prefix, CFG mix of the cond / uncond output halves
(``uncond + (cond - uncond) * cfg_scale``, generate.py:95-97), the mixed latent is both the output of the step and the
next input (through ``vae_latent_adapter``).  ``temperature`` / ``top_k`` / ``top_p`` / ``sample_logits`` are
accepted and ignored -- there is nothing to sample.  The whole loop runs inside libvlg.so (``vlg_generate_t2v``).
"""
import torch

from . import _lib as L


@torch.no_grad()
def generate(model, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, **sampling_kwargs):
    if model.model_type != 't2v':
        raise Exception("please check model type")
    T = cond.shape[1]
    assert T == model.cls_token_num
    dtype = model.tok_embeddings.weight.dtype
    cond_in = cond.to(dtype).contiguous()
    L.require_cuda(cond_in, "generate")
    B = cond.shape[0]
    rows = 2 * B if cfg_scale > 1.0 else B
    if emb_masks is not None:
        assert emb_masks.shape[0] == B and emb_masks.shape[-1] == T
    eng = model.engine(rows, T + max_new_tokens)
    model.max_batch_size, model.max_seq_length = rows, (T + max_new_tokens + 7) // 8 * 8
    return eng.generate_video(cond_in, max_new_tokens, emb_masks, cfg_scale, cfg_interval)
