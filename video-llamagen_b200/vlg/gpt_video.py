"""Drop-in for the inference surface of the reference's ``autoregressive/models/gpt_video.py`` (t2v trunk).

Same trunk as ``vlg.gpt`` (the reference file is a copy of gpt.py's blocks, gpt_video.py:147-267) with
  * the 3D RoPE table -- the 2D table tiled over the latent frames (gpt_video.py:532-552),
  * ``vae_latent_adapter`` (latent -> dim) feeding the trunk at decode time and ``vae_latent_adapter2`` (dim -> latent)
    on the way out (gpt_video.py:294-296, 410, 431): the model emits a continuous latent, there are no logits and no
    sampler,
  * a 120-token caption prefix with ``emb_masks`` like t2i.
Only ``shuffle_video_tokens=False`` is meaningful at inference (the reference's shuffled branch dereferences an
undefined ``ids_restore``).  Training branches raise.
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from .gpt import (CaptionEmbedder, LabelEmbedder, MLP, RMSNorm, TransformerBlock, Engine, find_multiple,
                  precompute_freqs_cis_2d, precompute_freqs_cis_3d_video)


@dataclass
class ModelArgs:
    """gpt_video.py:29-60."""
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
    model_type: str = 't2v'
    vocab_size: int = 16384
    cls_token_num: int = 1
    block_size: int = 256
    max_batch_size: int = 32
    max_seq_len: int = 2048
    vae_embed_dim: int = 2048
    t_downsample_size: int = 4
    num_frames: int = 17
    shuffle_video_tokens: bool = True


def _pad_cols(t, mult):
    n = (t.shape[-1] + mult - 1) // mult * mult
    if n == t.shape[-1]:
        return t.contiguous()
    out = torch.zeros(*t.shape[:-1], n, dtype=t.dtype, device=t.device)
    out[..., : t.shape[-1]] = t
    return out


class VideoEngine(Engine):
    """Engine plus the two adapter MLPs; in bf16 the 8-wide adapter ends are zero-padded to tensor-core granularity."""

    def __init__(self, model, max_rows, max_seq):
        super().__init__(model, max_rows, max_seq)

    def _extra_weights(self, model, w, dev):
        bf16 = self.dtype == torch.bfloat16
        fc1_in = model.vae_latent_adapter.fc1.weight.detach()
        fc2_out = model.vae_latent_adapter2.fc2.weight.detach()
        if bf16:
            fc1_in = _pad_cols(fc1_in, 64)                                     # [dim, 64]
            rows = (fc2_out.shape[0] + 15) // 16 * 16
            padded = torch.zeros(rows, fc2_out.shape[1], dtype=fc2_out.dtype, device=fc2_out.device)
            padded[: fc2_out.shape[0]] = fc2_out
            fc2_out = padded                                                   # [16, dim]
        w.adapter_in_fc1 = dev(fc1_in)
        w.adapter_in_fc2 = dev(model.vae_latent_adapter.fc2.weight)
        w.adapter_out_fc1 = dev(model.vae_latent_adapter2.fc1.weight)
        w.adapter_out_fc2 = dev(fc2_out)

    def generate_video(self, cond, max_new_tokens, emb_masks, cfg_scale, cfg_interval):
        """Whole t2v loop inside libvlg.so -> latents [B, max_new_tokens, vae_embed_dim] (model dtype)."""
        import ctypes as C
        B = cond.shape[0]
        gp = L.GenerateParams(B, int(max_new_tokens), float(cfg_scale), int(cfg_interval), 1.0, 0, 1.0, 0, 0)
        out = torch.empty(B, max_new_tokens, self.geo.vae_embed_dim, dtype=self.dtype, device=self.device)
        masks = None
        if emb_masks is not None:
            masks = emb_masks.to(self.device).ne(0).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_generate_t2v(self.ctx, C.byref(gp), L.ptr(cond), L.ptr(self.uncond), L.ptr(masks), L.ptr(out),
                                             L.stream()))
        return out

    def adapter_in(self, latent):
        x = latent.to(self.dtype)
        if self.dtype == torch.bfloat16:
            x = _pad_cols(x, 64)
        return self.mlp(1, x, self.geo.dim)

    def adapter_out(self, hidden):
        v = self.geo.vae_embed_dim
        width = (v + 15) // 16 * 16 if self.dtype == torch.bfloat16 else v
        return self.mlp(2, hidden, width)[..., :v]


class Transformer(nn.Module):
    """gpt_video.py:270-467, inference branch (sub-module order as in the reference so seeds and keys line up)."""

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
        head_dim = config.dim // config.n_head
        if self.model_type == 'c2i':
            self.cls_embedding = LabelEmbedder(config.num_classes, config.dim, config.class_dropout_prob)
        elif self.model_type == 't2i':
            self.cls_embedding = CaptionEmbedder(config.caption_dim, config.dim, config.class_dropout_prob)
            self.freqs_cis = precompute_freqs_cis_2d(grid_size, head_dim, config.rope_base, self.cls_token_num)
        elif self.model_type == 't2v':
            self.cls_embedding = CaptionEmbedder(config.caption_dim, config.dim, config.class_dropout_prob)
            self.vae_latent_adapter = MLP(config.vae_embed_dim, config.dim, config.dim)
            self.freqs_cis = precompute_freqs_cis_3d_video(grid_size, self.vae_t, head_dim, config.rope_base,
                                                           self.cls_token_num)
            self.vae_latent_adapter2 = MLP(config.dim, config.dim, config.vae_embed_dim)
            self.shuffle_video_tokens = config.shuffle_video_tokens
        else:
            raise Exception("please check model type")
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

    @property
    def vae_t(self):
        return (self.num_frames - 1) // self.t_downsample_size + 1

    def initialize_weights(self):
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
            self._engine = VideoEngine(self, rows, seq)
        return self._engine

    def setup_caches(self, max_batch_size, max_seq_length, dtype=None):
        """gpt_video.py:336-357."""
        self.max_seq_length = find_multiple(max_seq_length, 8)
        self.max_batch_size = max_batch_size
        self.engine(max_batch_size, self.max_seq_length).begin(max_batch_size)

    def set_prefix_mask(self, emb_masks_rows):
        self._engine.begin(self.max_batch_size, emb_masks_rows)

    def forward(self, idx=None, cond_idx=None, input_pos=None, targets=None, mask=None, valid=None, cond_embed=None,
                video_latent=None, targets_video=None):
        """gpt_video.py:403-467 inference branch -> (latent [R, S, vae_embed_dim], None)."""
        if targets is not None or targets_video is not None or self.training:
            raise NotImplementedError("vlg implements the inference branch only (training is out of scope: DESIGN.md)")
        if self.model_type != 't2v':
            raise NotImplementedError("use vlg.gpt.Transformer for c2i / t2i")
        if getattr(self, "shuffle_video_tokens", False):
            raise NotImplementedError("shuffle_video_tokens=True has no working inference path in the reference "
                                      "(undefined ids_restore, gpt_video.py:435-439); construct with False")
        eng = self._engine
        if eng is None or eng.rows <= 0:
            raise L.VlgError("call setup_caches() before forward()")
        pos = [int(p) for p in input_pos.reshape(-1).tolist()]
        if cond_embed is not None and video_latent is None:                    # prefill over the caption prefix
            emb = eng.mlp(0, cond_embed.to(eng.dtype), self.config.dim)[:, :self.cls_token_num]
            hs = [eng.forward_pos(p, embeds=emb[:, i].contiguous(), want_logits=False, want_hidden=True)[1]
                  for i, p in enumerate(pos)]
            return eng.adapter_out(torch.stack(hs, dim=1)), None
        assert cond_embed is None and video_latent is not None and len(pos) == 1
        emb = eng.adapter_in(video_latent.reshape(video_latent.shape[0], -1))
        h = eng.forward_pos(pos[0], embeds=emb.contiguous(), want_logits=False, want_hidden=True)[1]
        return eng.adapter_out(h).unsqueeze(1), None


def _cfg(n_layer, n_head, dim):
    return lambda **kw: Transformer(ModelArgs(n_layer=n_layer, n_head=n_head, dim=dim, **kw))


GPT_models = {  # gpt_video.py:563-592
    'GPT-B': _cfg(12, 12, 768), 'GPT-L': _cfg(24, 16, 1024), 'GPT-XL': _cfg(36, 20, 1280), 'GPT-XXL': _cfg(48, 24, 1536),
    'GPT-XXXL': _cfg(48, 40, 2560), 'GPT-1B': _cfg(22, 32, 2048), 'GPT-3B': _cfg(24, 32, 3200), 'GPT-7B': _cfg(32, 32, 4096),
}
