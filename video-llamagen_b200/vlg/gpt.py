"""Drop-in for the reference's ``autoregressive/models/gpt.py`` (inference surface).

Same registry (``GPT_models``), same ``ModelArgs`` fields, same module tree and therefore the same
``state_dict`` keys and -- because parameters are created by the same torch constructors in the same order --
the same weights for a given ``torch.manual_seed`` (gpt.py:262-316).  The modules here only *hold* parameters:
every inference computation goes to libvlg.so through :class:`Engine`.  Training branches of the reference
``forward`` (targets / loss, gpt.py:343-347,373-383) are out of scope and raise.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, List

import torch
import torch.nn as nn

from . import _lib as L


def find_multiple(n: int, k: int):
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class ModelArgs:
    """Field-for-field the reference's ModelArgs (gpt.py:23-50)."""
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
    model_type: str = 'c2i'

    vocab_size: int = 16384
    cls_token_num: int = 1
    block_size: int = 256
    max_batch_size: int = 32
    max_seq_len: int = 2048


# ------------------------------------------------------------------------------- parameter holders
class LabelEmbedder(nn.Module):
    """gpt.py:56-83.  Table has num_classes + 1 rows; the extra row is the CFG null class."""

    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob


class MLP(nn.Module):
    """gpt.py:118-131: fc1 -> GELU(tanh) -> fc2, no bias."""

    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features, bias=False)
        self.act = nn.GELU(approximate='tanh')
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features, bias=False)


class CaptionEmbedder(nn.Module):
    """gpt.py:89-115.  uncond_embedding is always [120, in_channels] (token_num default)."""

    def __init__(self, in_channels, hidden_size, uncond_prob, token_num=120):
        super().__init__()
        self.cap_proj = MLP(in_channels, hidden_size, hidden_size)
        self.register_buffer("uncond_embedding", nn.Parameter(torch.randn(token_num, in_channels) / in_channels ** 0.5))
        self.uncond_prob = uncond_prob


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        """gpt.py:143-148 on the device kernel (two roundings: normalise, then scale)."""
        L.require_cuda(x, "RMSNorm")
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        out = torch.empty_like(x2)
        L.check(L.lib().vlg_rmsnorm(L.ptr(x2), L.ptr(self.weight.detach().contiguous()), L.ptr(out), x2.shape[0],
                                    x2.shape[1], self.eps, L.dtype_code(x.dtype), L.stream()))
        return out.view(x.shape)


class FeedForward(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        hidden_dim = int(2 * (4 * config.dim) / 3)
        if config.ffn_dim_multiplier is not None:
            hidden_dim = int(config.ffn_dim_multiplier * hidden_dim)
        hidden_dim = find_multiple(hidden_dim, config.multiple_of)
        self.w1 = nn.Linear(config.dim, hidden_dim, bias=False)
        self.w3 = nn.Linear(config.dim, hidden_dim, bias=False)
        self.w2 = nn.Linear(hidden_dim, config.dim, bias=False)
        self.ffn_dropout = nn.Dropout(config.ffn_dropout_p)


class Attention(nn.Module):
    def __init__(self, config: ModelArgs):
        super().__init__()
        assert config.dim % config.n_head == 0
        self.dim = config.dim
        self.head_dim = config.dim // config.n_head
        self.n_head = config.n_head
        self.n_kv_head = config.n_kv_head if config.n_kv_head is not None else config.n_head
        if self.n_kv_head != self.n_head:
            raise NotImplementedError("grouped-query attention: no shipped reference config sets n_kv_head (gpt.py:195)")
        self.wqkv = nn.Linear(config.dim, (self.n_head + 2 * self.n_kv_head) * self.head_dim, bias=False)
        self.wo = nn.Linear(config.dim, config.dim, bias=False)
        self.kv_cache = None
        self.resid_dropout = nn.Dropout(config.resid_dropout_p)


class TransformerBlock(nn.Module):
    def __init__(self, config: ModelArgs, drop_path: float):
        super().__init__()
        self.attention = Attention(config)
        self.feed_forward = FeedForward(config)
        self.attention_norm = RMSNorm(config.dim, eps=config.norm_eps)
        self.ffn_norm = RMSNorm(config.dim, eps=config.norm_eps)
        self.drop_path = nn.Identity()


# ------------------------------------------------------------------------------- RoPE tables
def precompute_freqs_cis_2d(grid_size: int, n_elem: int, base: int = 10000, cls_token_num=120):
    """gpt.py:407-420: fp32 [cls_token_num + grid^2, n_elem/2, 2]; first half of the pairs turns with the row,
    second half with the column; conditioning rows are all-zero."""
    half = n_elem // 2
    inv = 1.0 / (base ** (torch.arange(0, half, 2)[: (half // 2)].float() / half))
    ang = torch.outer(torch.arange(grid_size), inv)
    grid = torch.concat([ang[:, None, :].expand(-1, grid_size, -1), ang[None, :, :].expand(grid_size, -1, -1)], dim=-1)
    cache = torch.stack([torch.cos(grid), torch.sin(grid)], dim=-1).flatten(0, 1)
    return torch.cat([torch.zeros(cls_token_num, half, 2), cache])


def precompute_freqs_cis_3d_video(grid_size: int, vae_t, n_elem: int, base: int = 10000, cls_token_num=120):
    """gpt_video.py:532-552: the 2D table repeated vae_t times."""
    cache = precompute_freqs_cis_2d(grid_size, n_elem, base, 0)
    return torch.cat([torch.zeros(cls_token_num, n_elem // 2, 2), cache.repeat(vae_t, 1, 1)])


# ------------------------------------------------------------------------------- engine
class Engine:
    """Owns one ``vlg_ctx``: packed weights, workspace (KV cache + activations) and the decode graph."""

    def __init__(self, model, max_rows, max_seq):
        cfg = model.config
        p = next(model.parameters())
        if not p.is_cuda:
            raise L.VlgError("model is on CPU: the vlg decode path needs a B200 (no CPU fallback)")
        self.device, self.dtype = p.device, p.dtype
        self.max_rows, self.max_seq = int(max_rows), int(max_seq)
        dt = L.dtype_code(self.dtype)
        mt = L.MODEL_TYPES[model.model_type]
        hid = model.layers[0].feed_forward.w1.weight.shape[0]
        self.geo = L.Geometry(cfg.dim, cfg.n_layer, cfg.n_head, hid, cfg.vocab_size, model.cls_token_num,
                              cfg.num_classes, cfg.caption_dim, getattr(cfg, "vae_embed_dim", 0), mt, dt, cfg.norm_eps)
        self.keep = []  # tensors the C side borrows

        def dev(t):
            t = t.detach().to(device=self.device, dtype=self.dtype).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        layers = (L.LayerWeights * cfg.n_layer)()
        for i, blk in enumerate(model.layers):
            w1, w3 = blk.feed_forward.w1.weight.detach(), blk.feed_forward.w3.weight.detach()
            w13 = torch.stack([w1, w3], dim=1).reshape(2 * hid, cfg.dim)      # row 2j = w1[j], 2j+1 = w3[j]
            layers[i] = L.LayerWeights(dev(blk.attention_norm.weight), dev(blk.attention.wqkv.weight),
                                       dev(blk.attention.wo.weight), dev(blk.ffn_norm.weight), dev(w13),
                                       dev(blk.feed_forward.w2.weight))
        self.layers = layers
        table = model.freqs_cis.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(table)
        w = L.Weights()
        w.tok_embeddings = dev(model.tok_embeddings.weight)
        if model.model_type == 'c2i':
            w.cls_table = dev(model.cls_embedding.embedding_table.weight)
        elif isinstance(model.cls_embedding, CaptionEmbedder):
            w.cap_fc1 = dev(model.cls_embedding.cap_proj.fc1.weight)
            w.cap_fc2 = dev(model.cls_embedding.cap_proj.fc2.weight)
            self.uncond = model.cls_embedding.uncond_embedding.detach().to(self.device, self.dtype).contiguous()
        self._extra_weights(model, w, dev)
        w.norm = dev(model.norm.weight)
        w.output = dev(model.output.weight)
        w.rope_table = table.data_ptr()
        w.rope_rows = table.shape[0]
        w.layers = layers
        self.weights = w
        lib = L.lib()
        nbytes = lib.vlg_ctx_workspace_bytes(C.byref(self.geo), self.max_rows, self.max_seq)
        self.workspace = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
        base = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        self.ctx = C.c_void_p()
        with torch.cuda.device(self.device):
            L.check(lib.vlg_ctx_create(C.byref(self.geo), C.byref(w), self.max_rows, self.max_seq, C.c_void_p(base),
                                       nbytes, C.byref(self.ctx)))
        self.rows = 0

    def _extra_weights(self, model, w, dev):
        """Hook for model variants with more projections (vlg.gpt_video)."""

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                L.lib().vlg_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    def check(self):
        """Synchronise and raise if a decode launch since the last check lost its SMs to other work (vlg_ctx_check)."""
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_ctx_check(self.ctx, L.stream()))

    def begin(self, rows, prefix_mask=None):
        m = None
        if prefix_mask is not None:
            m = prefix_mask.to(device=self.device).ne(0).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_begin(self.ctx, int(rows), L.ptr(m), L.stream()))
        self.rows = int(rows)

    def forward_pos(self, pos, tokens=None, labels=None, embeds=None, want_logits=True, want_hidden=False):
        R = self.rows
        logits = torch.empty(R, self.geo.vocab_size, dtype=torch.float32, device=self.device) if want_logits else None
        hidden = torch.empty(R, self.geo.dim, dtype=self.dtype, device=self.device) if want_hidden else None
        for t in (tokens, labels):
            if t is not None:
                assert t.dtype == torch.int64 and t.numel() == R
        if embeds is not None:
            assert embeds.dtype == self.dtype and embeds.shape == (R, self.geo.dim)
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_forward_pos(self.ctx, L.ptr(tokens), L.ptr(labels), L.ptr(embeds), int(pos),
                                            L.ptr(logits), L.ptr(hidden), L.stream()))
        return logits, hidden

    def forward_prefix(self, embeds, want_logits=False, want_hidden=True):
        """All cls_token_num prefix positions at once -> (logits, hidden) of the last one (vlg_forward_prefix)."""
        R, T = self.rows, self.geo.cls_token_num
        assert embeds.dtype == self.dtype and embeds.shape == (R, T, self.geo.dim)
        logits = torch.empty(R, self.geo.vocab_size, dtype=torch.float32, device=self.device) if want_logits else None
        hidden = torch.empty(R, self.geo.dim, dtype=self.dtype, device=self.device) if want_hidden else None
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_forward_prefix(self.ctx, L.ptr(embeds.contiguous()), L.ptr(logits), L.ptr(hidden), L.stream()))
        return logits, hidden

    def mlp(self, which, x, out_features):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        out = torch.empty(x2.shape[0], out_features, dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_mlp(self.ctx, which, L.ptr(x2), L.ptr(out), x2.shape[0], L.stream()))
        return out.view(*x.shape[:-1], out_features)

    def generate(self, cond, max_new_tokens, emb_masks, cfg_scale, cfg_interval, temperature, top_k, top_p,
                 sample_logits, seed):
        B = cond.shape[0]
        gp = L.GenerateParams(B, int(max_new_tokens), float(cfg_scale), int(cfg_interval), float(temperature),
                              int(top_k or 0), float(top_p), int(bool(sample_logits)), int(seed))
        out = torch.empty(B, max_new_tokens, dtype=torch.int32, device=self.device)
        masks = None
        if emb_masks is not None:
            masks = emb_masks.to(self.device).ne(0).to(torch.uint8).contiguous()
        uncond = getattr(self, "uncond", None)
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_generate(self.ctx, C.byref(gp), L.ptr(cond), L.ptr(uncond), L.ptr(masks), L.ptr(out),
                                         L.stream()))
        return out

    def generate_host(self, labels_host, tokens_host, max_new_tokens, cfg_scale, cfg_interval, temperature, top_k,
                      top_p, sample_logits, seed):
        """c2i with HOST buffers end to end (the call bench.py's e2e leg times). Synchronises."""
        B = labels_host.shape[0]
        gp = L.GenerateParams(B, int(max_new_tokens), float(cfg_scale), int(cfg_interval), float(temperature),
                              int(top_k or 0), float(top_p), int(bool(sample_logits)), int(seed))
        assert labels_host.dtype == torch.int64 and not labels_host.is_cuda
        assert tokens_host.dtype == torch.int32 and tokens_host.shape == (B, max_new_tokens) and not tokens_host.is_cuda
        with torch.cuda.device(self.device):
            L.check(L.lib().vlg_generate_c2i_host(self.ctx, C.byref(gp), L.ptr(labels_host), L.ptr(tokens_host),
                                                  L.stream()))
        return tokens_host


# ------------------------------------------------------------------------------- model
class Transformer(nn.Module):
    """gpt.py:262-385 (inference).  Construction order of sub-modules matches the reference so that
    seeds and state_dict keys line up."""

    def __init__(self, config: ModelArgs):
        super().__init__()
        self.config = config
        self.vocab_size = config.vocab_size
        self.n_layer = config.n_layer
        self.block_size = config.block_size
        self.num_classes = config.num_classes
        self.model_type = config.model_type
        self.cls_token_num = config.cls_token_num
        if self.model_type == 'c2i':
            self.cls_embedding = LabelEmbedder(config.num_classes, config.dim, config.class_dropout_prob)
        elif self.model_type == 't2i':
            self.cls_embedding = CaptionEmbedder(config.caption_dim, config.dim, config.class_dropout_prob)
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

        grid_size = int(self.block_size ** 0.5)
        assert grid_size * grid_size == self.block_size
        self.freqs_cis = precompute_freqs_cis_2d(grid_size, config.dim // config.n_head, config.rope_base,
                                                 self.cls_token_num)
        self.max_batch_size = -1
        self.max_seq_length = -1
        self._engine = None
        self.initialize_weights()

    # -- init (gpt.py:302-316)
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

    # -- anything that can change weights / device / dtype invalidates the packed engine
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
            e = None
            torch.cuda.empty_cache()
            self._engine = Engine(self, rows, seq)
        return self._engine

    # -- gpt.py:318-332
    def setup_caches(self, max_batch_size, max_seq_length, dtype=None):
        self.max_seq_length = find_multiple(max_seq_length, 8)
        self.max_batch_size = max_batch_size
        self.engine(max_batch_size, self.max_seq_length).begin(max_batch_size)
        self._prefix_mask = None

    def set_prefix_mask(self, emb_masks_rows):
        """emb_masks for all cache rows (generate.py:156-165), [rows, cls_token_num], nonzero = visible."""
        self._engine.begin(self.max_batch_size, emb_masks_rows)

    # -- gpt.py:334-385, inference branch only
    def forward(self, idx, cond_idx, input_pos=None, targets=None, mask=None, valid=None):
        if (idx is not None and cond_idx is not None) or targets is not None or self.training:
            raise NotImplementedError("vlg implements the inference branch of Transformer.forward only "
                                      "(training / loss are out of scope: DESIGN.md)")
        if self._engine is None or self._engine.rows <= 0:
            raise L.VlgError("call setup_caches() before forward()")
        eng = self._engine
        pos = [int(p) for p in input_pos.reshape(-1).tolist()]
        if cond_idx is not None:                                              # prefill
            if self.model_type == 'c2i':
                assert pos == [0]
                logits, _ = eng.forward_pos(0, labels=cond_idx.reshape(-1).to(torch.int64).contiguous())
                return logits.unsqueeze(1), None
            emb = eng.mlp(0, cond_idx.to(eng.dtype), self.config.dim)[:, :self.cls_token_num]
            outs = [eng.forward_pos(p, embeds=emb[:, i].contiguous())[0] for i, p in enumerate(pos)]
            return torch.stack(outs, dim=1), None
        assert idx.shape[1] == 1 and len(pos) == 1                            # decode, seqlen 1
        logits, _ = eng.forward_pos(pos[0], tokens=idx.reshape(-1).to(torch.int64).contiguous())
        return logits.unsqueeze(1), None

    def get_fsdp_wrap_module_list(self) -> List[nn.Module]:
        return list(self.layers)


# ------------------------------------------------------------------------------- configs (gpt.py:441-470)
# All eight names construct (same parameters, same state_dict keys as the reference).  The decode kernels are written for
# head_dim = 64 -- GPT-B / L / XL / XXL / XXXL / 1B: every BASELINE config; GPT-3B (head_dim 100) and GPT-7B (128) build and
# load checkpoints but `vlg_ctx_create` refuses them with VLG_ERR_ARG at the first setup_caches() / forward.
def GPT_7B(**kwargs):
    return Transformer(ModelArgs(n_layer=32, n_head=32, dim=4096, **kwargs))

def GPT_3B(**kwargs):
    return Transformer(ModelArgs(n_layer=24, n_head=32, dim=3200, **kwargs))

def GPT_1B(**kwargs):
    return Transformer(ModelArgs(n_layer=22, n_head=32, dim=2048, **kwargs))

def GPT_XXXL(**kwargs):
    return Transformer(ModelArgs(n_layer=48, n_head=40, dim=2560, **kwargs))

def GPT_XXL(**kwargs):
    return Transformer(ModelArgs(n_layer=48, n_head=24, dim=1536, **kwargs))

def GPT_XL(**kwargs):
    return Transformer(ModelArgs(n_layer=36, n_head=20, dim=1280, **kwargs))

def GPT_L(**kwargs):
    return Transformer(ModelArgs(n_layer=24, n_head=16, dim=1024, **kwargs))

def GPT_B(**kwargs):
    return Transformer(ModelArgs(n_layer=12, n_head=12, dim=768, **kwargs))


GPT_models = {
    'GPT-B': GPT_B, 'GPT-L': GPT_L, 'GPT-XL': GPT_XL, 'GPT-XXL': GPT_XXL, 'GPT-XXXL': GPT_XXXL,
    'GPT-1B': GPT_1B, 'GPT-3B': GPT_3B, 'GPT-7B': GPT_7B,
}
