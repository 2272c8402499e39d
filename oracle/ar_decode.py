"""TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's autoregressive
decode path (autoregressive/models/gpt.py + generate.py, and the t2v trunk of
gpt_video.py).  See oracle/__init__.py for who may import this and how it is pinned.

It is written functionally over a plain ``dict`` of weights that uses the
reference's ``state_dict`` key names, so a test can hand the *same* tensors to the
reference module, to this oracle and to the CUDA product.  Every function cites the
reference lines it restates.  It runs on CPU (fp32 contract, CPU baseline timing)
and, on the GPU box, on ``cuda`` through torch library kernels as the bf16 checker
(the bf16 rounding points are only reproducible on the device).
"""
import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- config

def find_multiple(n, k):
    """gpt.py:18-21."""
    return n if n % k == 0 else n + k - (n % k)


@dataclass
class Geometry:
    """The subset of gpt.py:23-50 ``ModelArgs`` that shapes the decode step."""
    dim: int
    n_layer: int
    n_head: int
    vocab_size: int = 16384
    block_size: int = 256
    cls_token_num: int = 1
    num_classes: int = 1000
    caption_dim: int = 2048
    model_type: str = "c2i"
    rope_base: float = 10000.0
    norm_eps: float = 1e-5
    multiple_of: int = 256
    # t2v only (gpt_video.py ModelArgs)
    vae_embed_dim: int = 8
    num_frames: int = 17
    t_downsample_size: int = 4

    @property
    def head_dim(self):
        return self.dim // self.n_head

    @property
    def hidden_dim(self):
        """gpt.py:154-159."""
        return find_multiple(int(2 * (4 * self.dim) / 3), self.multiple_of)

    @property
    def vae_t(self):
        """gpt_video.py:296 / sample_t2v_1f.py:198."""
        return (self.num_frames - 1) // self.t_downsample_size + 1


NAMED = {  # gpt.py:441-470
    "GPT-B": dict(n_layer=12, n_head=12, dim=768),
    "GPT-L": dict(n_layer=24, n_head=16, dim=1024),
    "GPT-XL": dict(n_layer=36, n_head=20, dim=1280),
    "GPT-XXL": dict(n_layer=48, n_head=24, dim=1536),
    "GPT-XXXL": dict(n_layer=48, n_head=40, dim=2560),
    "GPT-1B": dict(n_layer=22, n_head=32, dim=2048),
    "GPT-3B": dict(n_layer=24, n_head=32, dim=3200),
    "GPT-7B": dict(n_layer=32, n_head=32, dim=4096),
}


def init_weights(geo: Geometry, seed=0, dtype=torch.float32, output_std=0.02, embed_std=0.02):
    """Seeded random weights under the reference's state_dict names (gpt.py:262-316).

    Not a replay of the reference's RNG order -- fixtures carry their own weights or
    are generated from the reference module directly.  ``output.weight`` is drawn
    from normal(0, output_std) because the reference zero-initialises it
    (gpt.py:307), which would make every logit 0 and every parity test vacuous.
    """
    g = torch.Generator().manual_seed(seed)

    def n(*shape, std=0.02):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    w = {}
    d, hid = geo.dim, geo.hidden_dim
    if geo.model_type == "c2i":
        w["cls_embedding.embedding_table.weight"] = n(geo.num_classes + 1, d, std=embed_std)
    else:
        w["cls_embedding.cap_proj.fc1.weight"] = n(d, geo.caption_dim)
        w["cls_embedding.cap_proj.fc2.weight"] = n(d, d)
        w["cls_embedding.uncond_embedding"] = (
            torch.randn(120, geo.caption_dim, generator=g) / geo.caption_dim ** 0.5).to(dtype)
    if geo.model_type == "t2v":
        w["vae_latent_adapter.fc1.weight"] = n(d, geo.vae_embed_dim)
        w["vae_latent_adapter.fc2.weight"] = n(d, d)
        w["vae_latent_adapter2.fc1.weight"] = n(d, d)
        w["vae_latent_adapter2.fc2.weight"] = n(geo.vae_embed_dim, d)
    w["tok_embeddings.weight"] = n(geo.vocab_size, d, std=embed_std)
    for i in range(geo.n_layer):
        p = f"layers.{i}."
        w[p + "attention.wqkv.weight"] = n(3 * d, d)
        w[p + "attention.wo.weight"] = n(d, d)
        w[p + "feed_forward.w1.weight"] = n(hid, d)
        w[p + "feed_forward.w3.weight"] = n(hid, d)
        w[p + "feed_forward.w2.weight"] = n(d, hid)
        w[p + "attention_norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g)).to(dtype)
        w[p + "ffn_norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    w["norm.weight"] = (1.0 + 0.1 * torch.randn(d, generator=g)).to(dtype)
    w["output.weight"] = n(geo.vocab_size, d, std=output_std)
    return w


# --------------------------------------------------------------------------- RoPE tables

def rope_table_2d(grid_size, n_elem, base=10000.0, cls_token_num=1):
    """gpt.py:407-420.  fp32 [cls_token_num + grid^2, n_elem/2, 2] = (cos, sin).

    Pair k < n_elem/4 rotates with the row coordinate, pair k >= n_elem/4 with the
    column coordinate; the cls_token_num leading rows are all-zero (cos = sin = 0),
    so q and k of conditioning tokens become exactly 0.
    """
    half = n_elem // 2
    inv = 1.0 / (base ** (torch.arange(0, half, 2)[: half // 2].float() / half))
    ang = torch.outer(torch.arange(grid_size), inv)                  # [grid, half/2]
    ang_grid = torch.concat([ang[:, None, :].expand(-1, grid_size, -1),
                             ang[None, :, :].expand(grid_size, -1, -1)], dim=-1)
    tab = torch.stack([torch.cos(ang_grid), torch.sin(ang_grid)], dim=-1).flatten(0, 1)
    return torch.cat([torch.zeros(cls_token_num, half, 2), tab])


def rope_table_3d(grid_size, vae_t, n_elem, base=10000.0, cls_token_num=120):
    """gpt_video.py:532-552: the 2D table tiled vae_t times (no temporal frequency)."""
    tab2d = rope_table_2d(grid_size, n_elem, base, 0)
    return torch.cat([torch.zeros(cls_token_num, n_elem // 2, 2), tab2d.repeat(vae_t, 1, 1)])


def rope_table(geo: Geometry):
    grid = int(geo.block_size ** 0.5)
    assert grid * grid == geo.block_size            # gpt.py:292-293
    if geo.model_type == "t2v":
        return rope_table_3d(grid, geo.vae_t, geo.head_dim, geo.rope_base, geo.cls_token_num)
    return rope_table_2d(grid, geo.head_dim, geo.rope_base, geo.cls_token_num)


def rotate_pairs(x, rows):
    """gpt.py:423-433.  x [B,S,H,D]; rows [S,D/2,2] fp32.  fp32 math, each product
    rounded separately, result cast back to x.dtype."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    t = rows.view(1, xs.size(1), 1, xs.size(3), 2)
    out = torch.stack([xs[..., 0] * t[..., 0] - xs[..., 1] * t[..., 1],
                       xs[..., 1] * t[..., 0] + xs[..., 0] * t[..., 1]], dim=-1)
    return out.flatten(3).type_as(x)


# --------------------------------------------------------------------------- blocks

def rms_norm(x, weight, eps):
    """gpt.py:143-148: fp32 normalise, round to x.dtype, then multiply by weight."""
    xf = x.float()
    y = xf * torch.rsqrt(torch.mean(xf * xf, dim=-1, keepdim=True) + eps)
    return y.type_as(x) * weight


def sdpa_math(q, k, v, mask):
    """torch's math SDPA backend, which generate.py:112 forces for every decode step
    (SURVEY.md 3.5 step 6): fp32 upcast, q and k each scaled by sqrt(1/sqrt(D)),
    boolean mask -> additive -inf, fp32 softmax, fp32 P.V, one rounding at the end."""
    s = math.sqrt(1.0 / math.sqrt(q.size(-1)))
    qf, kf, vf = q.float() * s, k.float() * s, v.float()
    att = qf @ kf.transpose(-2, -1)
    att = att.masked_fill(~mask, float("-inf"))
    att = torch.softmax(att, dim=-1)
    return (att @ vf).to(q.dtype)


def mlp_gelu_tanh(x, fc1, fc2):
    """gpt.py:118-131 (bias-free fc1 -> GELU(tanh) -> fc2)."""
    return F.linear(F.gelu(F.linear(x, fc1), approximate="tanh"), fc2)


class DecodeState:
    """gpt.py:318-332 ``setup_caches`` + generate.py:156-165 mask editing."""

    def __init__(self, geo: Geometry, rows, max_seq_length, dtype, device="cpu"):
        self.S = find_multiple(max_seq_length, 8)
        shape = (rows, geo.n_head, self.S, geo.head_dim)
        self.k = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(geo.n_layer)]
        self.v = [torch.zeros(shape, dtype=dtype, device=device) for _ in range(geo.n_layer)]
        tri = torch.tril(torch.ones(self.S, self.S, dtype=torch.bool, device=device))
        self.mask = tri.unsqueeze(0).repeat(rows, 1, 1)
        self.table = rope_table(geo).to(device)

    def apply_prefix_mask(self, emb_masks_rows):
        """generate.py:156-165: hide padded prefix keys from every query, then force
        the diagonal so a padded query still attends to itself."""
        T = emb_masks_rows.shape[-1]
        self.mask[:, :, :T] = self.mask[:, :, :T] * emb_masks_rows.to(self.mask.device).unsqueeze(1).bool()
        eye = torch.eye(self.S, dtype=torch.bool, device=self.mask.device)
        self.mask |= eye


def trunk(geo: Geometry, w, h, input_pos, st: DecodeState, faithful_copy=True):
    """gpt.py:356-370 + 207-259: N transformer blocks over h [R,S,d] at positions
    ``input_pos`` [S], updating the KV caches in ``st``.  Returns final-normed h."""
    R, S, d = h.shape
    H, D = geo.n_head, geo.head_dim
    mask = st.mask[:R, None, input_pos]                              # [R,1,S,S_max]
    rows = st.table[input_pos]
    for i in range(geo.n_layer):
        p = f"layers.{i}."
        xn = rms_norm(h, w[p + "attention_norm.weight"], geo.norm_eps)
        q, k, v = F.linear(xn, w[p + "attention.wqkv.weight"]).split([d, d, d], dim=-1)
        q = rotate_pairs(q.view(R, S, H, D), rows).transpose(1, 2)
        k = rotate_pairs(k.view(R, S, H, D), rows).transpose(1, 2)
        v = v.view(R, S, H, D).transpose(1, 2)
        st.k[i][:, :, input_pos] = k                                 # gpt.py:182-183
        st.v[i][:, :, input_pos] = v
        keys, values = st.k[i][:R], st.v[i][:R]
        if faithful_copy:                                            # gpt.py:230-231
            keys, values = keys.repeat_interleave(1, dim=1), values.repeat_interleave(1, dim=1)
        a = sdpa_math(q, keys, values, mask)                         # over all S_max slots
        a = a.transpose(1, 2).contiguous().view(R, S, d)
        h = h + F.linear(a, w[p + "attention.wo.weight"])            # gpt.py:241,257
        xn = rms_norm(h, w[p + "ffn_norm.weight"], geo.norm_eps)
        g = F.silu(F.linear(xn, w[p + "feed_forward.w1.weight"])) * F.linear(xn, w[p + "feed_forward.w3.weight"])
        h = h + F.linear(g, w[p + "feed_forward.w2.weight"])         # gpt.py:167,258
    return rms_norm(h, w["norm.weight"], geo.norm_eps)


def embed_condition(geo, w, cond):
    """gpt.py:78-83 (c2i label gather) / gpt.py:110-131 (t2i,t2v caption MLP)."""
    if geo.model_type == "c2i":
        return w["cls_embedding.embedding_table.weight"][cond].unsqueeze(1)[:, :geo.cls_token_num]
    return mlp_gelu_tanh(cond, w["cls_embedding.cap_proj.fc1.weight"],
                         w["cls_embedding.cap_proj.fc2.weight"])[:, :geo.cls_token_num]


def forward_logits(geo, w, st, idx=None, cond=None, input_pos=None):
    """gpt.py:349-371 inference branch: fp32 logits [R,S,V]."""
    h = embed_condition(geo, w, cond) if cond is not None else w["tok_embeddings.weight"][idx]
    h = trunk(geo, w, h, input_pos, st)
    return F.linear(h, w["output.weight"]).float()


def forward_latent(geo, w, st, latent=None, cond=None, input_pos=None):
    """gpt_video.py:403-431 inference branch (shuffle_video_tokens=False): the trunk
    between the two adapter MLPs; output is a continuous latent [R,S,vae_embed_dim]."""
    if cond is not None:
        h = embed_condition(geo, w, cond)
    else:
        h = mlp_gelu_tanh(latent, w["vae_latent_adapter.fc1.weight"], w["vae_latent_adapter.fc2.weight"])
    h = trunk(geo, w, h, input_pos, st)
    return mlp_gelu_tanh(h, w["vae_latent_adapter2.fc1.weight"], w["vae_latent_adapter2.fc2.weight"])


# --------------------------------------------------------------------------- sampling

def cfg_mix(logits, cfg_scale):
    """generate.py:81-82 / 95-97: rows [0,B) conditional, [B,2B) unconditional."""
    c, u = torch.split(logits, len(logits) // 2, dim=0)
    return u + (c - u) * cfg_scale


def top_k_top_p_filter(logits, top_k=0, top_p=1.0):
    """generate.py:16-54, in place on logits [B,V].

    top-k: threshold = k-th largest value; entries strictly below it become -inf
    (ties at the threshold all survive).  top-p: sort descending, softmax, cumsum;
    drop entries whose cumulative mass exceeds top_p, shifted right by one so the
    first entry that crosses is kept; scatter back."""
    if top_k > 0:
        k = min(max(top_k, 1), logits.size(-1))
        kth = torch.topk(logits, k)[0][..., -1, None]
        logits[logits < kth] = float("-inf")
    if top_p < 1.0:
        s_logits, s_idx = torch.sort(logits, descending=True)
        drop = torch.cumsum(torch.softmax(s_logits, dim=-1), dim=-1) > top_p
        drop[..., 1:] = drop[..., :-1].clone()
        drop[..., 0] = False
        logits[drop.scatter(1, s_idx, drop)] = float("-inf")
    return logits


def sample(logits, temperature=1.0, top_k=0, top_p=1.0, sample_logits=True, noise=None):
    """generate.py:57-66.  logits [B,S,V] -> (idx int64 [B,1], probs fp32 [B,V]).

    ``noise`` (optional, [B,V] Exp(1) draws) replaces the hidden RNG of
    ``torch.multinomial``, whose single-sample path is argmax(probs / Exp(1)); with it
    the sampled branch is a deterministic function that the CUDA sampler can be
    compared with draw for draw."""
    l = logits[:, -1, :] / max(temperature, 1e-5)
    if top_k > 0 or top_p < 1.0:
        l = top_k_top_p_filter(l, top_k=top_k, top_p=top_p)
    probs = torch.softmax(l, dim=-1)
    if not sample_logits:
        idx = torch.topk(probs, k=1, dim=-1)[1]
    elif noise is not None:
        idx = torch.argmax(probs / noise, dim=-1, keepdim=True)
    else:
        idx = torch.multinomial(probs, num_samples=1)
    return idx, probs


# --------------------------------------------------------------------------- generate

@torch.no_grad()
def generate(geo, w, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1,
             noise_fn=None, trace=None, **sampling):
    """generate.py:127-180 (+ prefill :77-86, decode_one_token :89-102,
    decode_n_tokens :105-124).  Returns int32 [B, max_new_tokens].

    ``noise_fn(step, B, V)`` optionally supplies the Exp(1) draws per step.
    ``trace`` (a list) receives the fp32 raw logits [R,V] of every step."""
    use_cfg = cfg_scale > 1.0
    dev = cond.device
    if geo.model_type == "c2i":
        rows = torch.cat([cond, torch.ones_like(cond) * geo.num_classes]) if use_cfg else cond
        T = 1
    else:
        null = torch.zeros_like(cond) + w["cls_embedding.uncond_embedding"]
        rows = torch.cat([cond, null]) if use_cfg else cond
        T = cond.shape[1]
    B = cond.shape[0]
    R = 2 * B if use_cfg else B
    st = DecodeState(geo, R, T + max_new_tokens, w["tok_embeddings.weight"].dtype, dev)
    if emb_masks is not None:
        st.apply_prefix_mask(torch.cat([emb_masks, emb_masks]) if use_cfg else emb_masks)
    seq = torch.empty((B, max_new_tokens), dtype=torch.int32, device=dev)

    def pick(logits, step, guided):
        if trace is not None:
            trace.append(logits[:, -1].clone())
        if use_cfg:
            logits = cfg_mix(logits, cfg_scale) if guided else logits[: len(logits) // 2]
        nz = noise_fn(step, B, logits.shape[-1]) if noise_fn is not None else None
        return sample(logits, noise=nz, **sampling)[0]

    tok = pick(forward_logits(geo, w, st, cond=rows, input_pos=torch.arange(T, device=dev)), 0, True)
    seq[:, 0:1] = tok
    guided = True
    for i in range(max_new_tokens - 1):
        if cfg_interval > -1 and i > cfg_interval:                   # generate.py:113-114
            guided = False
        x = torch.cat([tok, tok]) if use_cfg else tok                # generate.py:92
        pos = torch.tensor([T + i], device=dev)
        tok = pick(forward_logits(geo, w, st, idx=x.view(-1, 1), input_pos=pos), i + 1, guided)
        seq[:, i + 1:i + 2] = tok
    return seq


@torch.no_grad()
def generate_video(geo, w, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1):
    """The t2v loop the reference's samplers call (sample_t2v_1f.py:199-205; models/generate_video.py is missing from
    the fork): generate.py:127-180 with the token hand-off replaced by the latent hand-off of the gpt_video trunk
    (gpt_video.py:403-431).  Prefill over the caption prefix, CFG mix ``uncond + (cond - uncond) * scale`` of the two
    output halves (generate.py:95-97; the conditional half alone once ``i > cfg_interval``, :113-114), the mixed latent
    is the step's output and the next input.  Returns [B, max_new_tokens, vae_embed_dim] in the weights' dtype."""
    use_cfg = cfg_scale > 1.0
    dev = cond.device
    null = torch.zeros_like(cond) + w["cls_embedding.uncond_embedding"]
    rows = torch.cat([cond, null]) if use_cfg else cond
    T, B = cond.shape[1], cond.shape[0]
    R = 2 * B if use_cfg else B
    dt = w["tok_embeddings.weight"].dtype
    st = DecodeState(geo, R, T + max_new_tokens, dt, dev)
    if emb_masks is not None:
        st.apply_prefix_mask(torch.cat([emb_masks, emb_masks]) if use_cfg else emb_masks)

    def mix(out, guided):
        if not use_cfg:
            return out
        c, u = torch.split(out.float(), B, dim=0)
        return (u + (c - u) * cfg_scale if guided else c).to(dt)

    lat = mix(forward_latent(geo, w, st, cond=rows, input_pos=torch.arange(T, device=dev))[:, -1], True)
    seq = [lat]
    guided = True
    for i in range(max_new_tokens - 1):
        if cfg_interval > -1 and i > cfg_interval:
            guided = False
        x = (torch.cat([lat, lat]) if use_cfg else lat).unsqueeze(1)
        lat = mix(forward_latent(geo, w, st, latent=x, input_pos=torch.tensor([T + i], device=dev))[:, -1], guided)
        seq.append(lat)
    return torch.stack(seq, dim=1)


@torch.no_grad()
def teacher_forced_logits(geo, w, cond, tokens, emb_masks=None, cfg=True):
    """Per-step raw logits [N, R, V] when the token stream is fixed in advance
    (SURVEY.md 8c item 3): step 0 is the prefill's last position, step i>0 feeds
    tokens[:, i-1] at position T+i-1."""
    dev = cond.device
    if geo.model_type == "c2i":
        rows = torch.cat([cond, torch.ones_like(cond) * geo.num_classes]) if cfg else cond
        T = 1
    else:
        null = torch.zeros_like(cond) + w["cls_embedding.uncond_embedding"]
        rows = torch.cat([cond, null]) if cfg else cond
        T = cond.shape[1]
    N = tokens.shape[1] + 1
    R = rows.shape[0]
    st = DecodeState(geo, R, T + N, w["tok_embeddings.weight"].dtype, dev)
    if emb_masks is not None:
        st.apply_prefix_mask(torch.cat([emb_masks, emb_masks]) if cfg else emb_masks)
    out = [forward_logits(geo, w, st, cond=rows, input_pos=torch.arange(T, device=dev))[:, -1]]
    for i in range(N - 1):
        x = tokens[:, i:i + 1]
        x = torch.cat([x, x]) if cfg else x
        out.append(forward_logits(geo, w, st, idx=x, input_pos=torch.tensor([T + i], device=dev))[:, -1])
    return torch.stack(out)
