"""Drop-in for the inference surface of the reference's ``autoregressive/models/gpt_video_diff.py``: the t2v trunk whose
output head is a diffusion sampler.

Against ``vlg.gpt_video`` the model has no ``vae_latent_adapter2``: ``forward`` returns the final-normed hidden state
``[rows, S, dim]`` (gpt_video_diff.py:583-647, inference branch) and the caller draws the token's latent from
``model.diffloss.sample(hidden, temperature, cfg)`` (generate_video_diff.py:88-90, 130-132).  The trunk is the same
decode path as every other model here (``vlg_forward_pos`` with an embedding input, ``vlg_mlp`` for the caption and
latent adapters); the head is ``vlg.diffloss.DiffLoss`` (``vlg_diff_sample``).  Sub-modules are created in the
reference's order (gpt_video_diff.py:322-425), so a seed gives the same weights and ``state_dict`` keys match
(``mask_token``, ``cls_embedding.*``, ``vae_latent_adapter.*``, ``diffloss.net.*``, ``tok_embeddings``, ``layers.*``,
``norm``, ``output``).  Training and the masked-generation branches (``sample_tokens*``) raise.
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from .diffloss import DiffLoss
from .gpt import (CaptionEmbedder, MLP, RMSNorm, TransformerBlock, Engine, find_multiple, precompute_freqs_cis_3d_video)


@dataclass
class ModelArgs:
    """gpt_video_diff.py:42-82."""
    dim: int = 4096
    n_layer: int = 32
    n_head: int = 32
    n_kv_head: Optional[int] = None
    multiple_of: int = 256
    ffn_dim_multiplier: Optional[float] = None
    rope_base: float = 10000
    norm_eps: float = 1e-5
    initializer_range: float = 0.02
    token_dropout_p: float = 0.1
    attn_dropout_p: float = 0.0
    resid_dropout_p: float = 0.1
    ffn_dropout_p: float = 0.1
    drop_path_rate: float = 0.0
    num_classes: int = 1000
    caption_dim: int = 2048
    class_dropout_prob: float = 0.1
    model_type: str = "t2v"
    vocab_size: int = 16384
    cls_token_num: int = 1
    block_size: int = 256
    max_batch_size: int = 32
    max_seq_len: int = 2048
    vae_embed_dim: int = 2048
    t_downsample_size: int = 4
    num_frames: int = 17
    save_train_video_latent: bool = False
    diffloss_d: int = 3
    diffloss_w: int = 1024
    num_sampling_steps: int = "100"
    diffusion_batch_mul: int = 4
    grad_checkpointing: bool = False
    mask_ratio_min: float = 0.7


class DiffVideoEngine(Engine):
    """Engine plus the latent-in adapter (there is no latent-out adapter: the head is the diffusion sampler)."""

    def _extra_weights(self, model, w, dev):
        w.adapter_in_fc1 = dev(model.vae_latent_adapter.fc1.weight)
        w.adapter_in_fc2 = dev(model.vae_latent_adapter.fc2.weight)

    def adapter_in(self, latent):
        return self.mlp(1, latent.to(self.dtype), self.geo.dim)


class Transformer(nn.Module):
    """gpt_video_diff.py:321-425 + the inference branch of forward (:583-647)."""

    def __init__(self, config: ModelArgs):
        super().__init__()
        self.config = config
        self.vocab_size = config.vocab_size
        self.n_layer = config.n_layer
        self.block_size = config.block_size
        self.num_classes = config.num_classes
        self.model_type = config.model_type
        self.cls_token_num = config.cls_token_num
        self.num_frames = config.num_frames
        self.t_downsample_size = config.t_downsample_size
        grid_size = int(self.block_size ** 0.5)
        assert grid_size * grid_size == self.block_size
        if self.model_type != "t2v":
            raise NotImplementedError("vlg.gpt_video_diff covers the t2v model (the only one the fork's diffusion sampler "
                                      "drives); use vlg.gpt for c2i / t2i")
        self.cls_embedding = CaptionEmbedder(config.caption_dim, config.dim, config.class_dropout_prob)
        self.vae_latent_adapter = MLP(config.vae_embed_dim, config.dim, config.dim)
        self.vae_t = (config.num_frames - 1) // config.t_downsample_size + 1
        self.freqs_cis = precompute_freqs_cis_3d_video(grid_size, self.vae_t, config.dim // config.n_head,
                                                       config.rope_base, self.cls_token_num)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, config.vae_embed_dim))
        self.vae_embed_dim = config.vae_embed_dim
        self.grid_size = grid_size
        self.seq_len = int(grid_size ** 2 * self.vae_t)
        self.llm_dim = config.dim
        self.diffloss_d = config.diffloss_d
        self.diffloss_w = config.diffloss_w
        self.num_sampling_steps = config.num_sampling_steps
        self.diffusion_batch_mul = config.diffusion_batch_mul
        self.diffloss = DiffLoss(target_channels=config.vae_embed_dim, z_channels=config.dim, width=self.diffloss_w,
                                 depth=self.diffloss_d, num_sampling_steps=self.num_sampling_steps,
                                 grad_checkpointing=config.grad_checkpointing)
        self.tok_embeddings = nn.Embedding(config.vocab_size, config.dim)
        self.tok_dropout = nn.Dropout(config.token_dropout_p)
        dpr = [x.item() for x in torch.linspace(0, config.drop_path_rate, config.n_layer)]
        self.layers = torch.nn.ModuleList()
        for layer_id in range(config.n_layer):
            self.layers.append(TransformerBlock(config, dpr[layer_id]))
        self.norm = RMSNorm(config.dim, eps=config.norm_eps)
        self.output = nn.Linear(config.dim, config.vocab_size, bias=False)
        self.max_batch_size = -1
        self.max_seq_length = -1
        self._engine = None
        self.initialize_weights()

    def initialize_weights(self):
        """gpt_video_diff.py:429-444: every Linear / Embedding -- the diffusion head's included -- normal(0, 0.02), biases 0."""
        self.apply(self._init_weights)
        nn.init.constant_(self.output.weight, 0)

    def _init_weights(self, module):
        std = self.config.initializer_range
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=std)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=std)

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, *a, **k):
        self._engine = None
        return super().load_state_dict(state_dict, *a, **k)

    def engine(self, rows, seq):
        e = self._engine
        if e is None or e.max_rows < rows or e.max_seq < seq:
            self._engine = None
            torch.cuda.empty_cache()
            self._engine = DiffVideoEngine(self, rows, seq)
        return self._engine

    def setup_caches(self, max_batch_size, max_seq_length, dtype=None):
        """gpt_video_diff.py:446-472."""
        self.max_seq_length = find_multiple(max_seq_length, 8)
        self.max_batch_size = max_batch_size
        self.engine(max_batch_size, self.max_seq_length).begin(max_batch_size)

    def set_prefix_mask(self, emb_masks_rows):
        self._engine.begin(self.max_batch_size, emb_masks_rows)

    def forward(self, idx=None, cond_idx=None, input_pos=None, targets=None, mask=None, valid=None, cond_embed=None,
                video_latent=None, targets_video=None, mask_ratio=1, save_train_video_latent=False):
        """gpt_video_diff.py:583-647, inference branch -> (final-normed hidden [R, S, dim], None)."""
        if targets is not None or targets_video is not None or self.training:
            raise NotImplementedError("vlg implements the inference branch only (training is out of scope: DESIGN.md)")
        eng = self._engine
        if eng is None or eng.rows <= 0:
            raise L.VlgError("call setup_caches() before forward()")
        pos = [int(p) for p in input_pos.reshape(-1).tolist()]
        if cond_embed is not None and video_latent is None:                    # prefill over the caption prefix
            emb = eng.mlp(0, cond_embed.to(eng.dtype), self.config.dim)[:, :self.cls_token_num]
            hs = [eng.forward_pos(p, embeds=emb[:, i].contiguous(), want_logits=False, want_hidden=True)[1]
                  for i, p in enumerate(pos)]
            return torch.stack(hs, dim=1), None
        assert cond_embed is None and video_latent is not None and len(pos) == 1
        emb = eng.adapter_in(video_latent.reshape(video_latent.shape[0], -1))
        h = eng.forward_pos(pos[0], embeds=emb.contiguous(), want_logits=False, want_hidden=True)[1]
        return h.unsqueeze(1), None

    def sample_tokens(self, *a, **k):
        raise NotImplementedError("the masked-generation samplers of gpt_video_diff.py (sample_tokens*) are outside the decode "
                                  "hot path (DESIGN.md section 7)")

    sample_tokens2 = sample_tokens


def _cfg(n_layer, n_head, dim):
    return lambda **kw: Transformer(ModelArgs(n_layer=n_layer, n_head=n_head, dim=dim, **kw))


# gpt_video_diff.py:1036-1080.  head_dim 64 only: GPT-3B (100) and GPT-7B (128) construct but are rejected by vlg_ctx_create;
# GPT-1B (head_dim 64, the checkpoint the shipped script samples from) is the flagship of this path.
GPT_models = {
    'GPT-B': _cfg(12, 12, 768), 'GPT-L': _cfg(24, 16, 1024), 'GPT-XL': _cfg(36, 20, 1280), 'GPT-XXL': _cfg(48, 24, 1536),
    'GPT-XXXL': _cfg(48, 40, 2560), 'GPT-1B': _cfg(22, 32, 2048), 'GPT-3B': _cfg(24, 32, 3200), 'GPT-7B': _cfg(32, 32, 4096),
}
