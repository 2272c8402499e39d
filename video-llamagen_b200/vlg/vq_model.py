"""Drop-in for the quantizer of the reference's ``tokenizer/tokenizer_image/vq_model.py``.

``VectorQuantizer.forward`` (eval mode) and ``get_codebook_entry`` run on libvlg.so: fused normalise + distance +
running argmin with the codebook staged through shared memory -- the [N, n_e] distance matrix the reference
materialises (vq_model.py:228-230) never exists.  The VQGAN encoder / decoder CNNs are outside the hot path
(cuDNN convolutions; SURVEY.md 8f lists them as the next widening step).
"""
from dataclasses import dataclass, field
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L


@dataclass
class ModelArgs:
    """vq_model.py:12-24."""
    codebook_size: int = 16384
    codebook_embed_dim: int = 8
    codebook_l2_norm: bool = True
    codebook_show_usage: bool = True
    commit_loss_beta: float = 0.25
    entropy_loss_ratio: float = 0.0
    encoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    decoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    z_channels: int = 256
    dropout_p: float = 0.0


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim, beta, entropy_loss_ratio, l2_norm, show_usage):
        super().__init__()
        self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
        self.entropy_loss_ratio, self.l2_norm, self.show_usage = entropy_loss_ratio, l2_norm, show_usage
        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)
        if self.l2_norm:
            self.embedding.weight.data = F.normalize(self.embedding.weight.data, p=2, dim=-1)
        if self.show_usage:
            self.register_buffer("codebook_used", nn.Parameter(torch.zeros(65536)))

    @torch.no_grad()
    def forward(self, z):
        """vq_model.py:215-259, eval mode: z [B, C, H, W] fp32 ->
        (z_q [B, C, H, W], (None, None, None, 0), (None, None, indices int64 [B*H*W]))."""
        if self.training:
            raise NotImplementedError("vlg implements the eval-mode quantizer (losses / usage tracking are training-only)")
        L.require_cuda(z, "VectorQuantizer.forward")
        assert z.dim() == 4 and z.shape[1] == self.e_dim and z.dtype == torch.float32
        B, Cc, H, W = z.shape
        cb = self.embedding.weight.detach().to(torch.float32).contiguous()
        idx = torch.empty(B * H * W, dtype=torch.int64, device=z.device)
        z_q = torch.empty_like(z)
        scratch = torch.empty(L.lib().vlg_vq_scratch_floats(B * H * W, Cc, self.n_e), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            L.check(L.lib().vlg_vq_argmin(L.ptr(z), L.ptr(cb), L.ptr(idx), L.ptr(z_q), L.ptr(scratch), B, Cc, H * W,
                                          self.n_e, int(self.l2_norm), L.stream()))
        return z_q, (None, None, None, 0), (None, None, idx)

    @torch.no_grad()
    def get_codebook_entry(self, indices, shape=None, channel_first=True):
        """vq_model.py:261-276."""
        L.require_cuda(indices, "get_codebook_entry")
        cb = self.embedding.weight.detach().to(torch.float32).contiguous()
        flat = indices.reshape(-1).to(torch.int64).contiguous()
        n = flat.numel()
        if shape is not None and channel_first:
            B, Cc, H, W = shape
            assert B * H * W == n and Cc == self.e_dim
            out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=indices.device)
            args = (B, Cc, H * W, 1)
        else:
            out = torch.empty(n, self.e_dim, dtype=torch.float32, device=indices.device)
            args = (n, self.e_dim, 1, 0)
        with torch.cuda.device(indices.device):
            L.check(L.lib().vlg_codebook_gather(L.ptr(flat), L.ptr(cb), L.ptr(out), args[0], args[1], args[2], self.n_e,
                                                int(self.l2_norm), args[3], L.stream()))
        if shape is not None and not channel_first:
            out = out.view(shape)
        return out

    @torch.no_grad()
    def decode_front(self, indices, shape, conv_weight, conv_bias):
        """get_codebook_entry + the 1x1 ``post_quant_conv`` in one kernel (vq_model.py:47-55, 261-276): code indices
        [B*H*W] -> feature map [B, z_channels, H, W].  The ``[B, e_dim, H, W]`` quantised map is never written."""
        L.require_cuda(indices, "decode_front")
        B, Cc, H, W = shape
        flat = indices.reshape(-1).to(torch.int64).contiguous()
        assert flat.numel() == B * H * W and Cc == self.e_dim
        cb = self.embedding.weight.detach().to(torch.float32).contiguous()
        w = conv_weight.detach().to(torch.float32).reshape(conv_weight.shape[0], Cc).contiguous()
        b = conv_bias.detach().to(torch.float32).contiguous() if conv_bias is not None else None
        out = torch.empty(B, w.shape[0], H, W, dtype=torch.float32, device=indices.device)
        with torch.cuda.device(indices.device):
            L.check(L.lib().vlg_codebook_project(L.ptr(flat), L.ptr(cb), L.ptr(w), L.ptr(b), L.ptr(out), B, Cc, H * W,
                                                 self.n_e, w.shape[0], int(self.l2_norm), L.stream()))
        return out


# ----------------------------------------------------------------------------------------------- tokenizer CNNs
# The VQGAN encoder / decoder around the quantizer (vq_model.py:65-194, 279-385).  The module tree -- and therefore every
# state_dict key and, parameters being created in the same order, every weight for a given seed -- is the reference's:
#   {encoder,decoder}.conv_in, .mid.{0,1,2}, .conv_blocks.<level>.res.<i>.{norm1,conv1,norm2,conv2,nin_shortcut},
#   .conv_blocks.<level>.attn.<i>.{norm,q,k,v,proj_out}, .conv_blocks.<level>.{downsample,upsample}.conv, .norm_out, .conv_out
# Convolutions, GroupNorm and the single-head spatial attention are library calls (cuDNN / cuBLAS through torch): they are
# conv-bound and outside the decode hot path; the hand-written parts are the index search in front of them and the fused
# codebook gather + post_quant_conv behind it.
def _gn(ch):
    return nn.GroupNorm(num_groups=32, num_channels=ch, eps=1e-6, affine=True)


def _swish(x):
    return F.silu(x)


class ResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels=None, dropout=0.0):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _gn(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _gn(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, x):
        y = self.conv1(_swish(self.norm1(x)))
        y = self.conv2(self.dropout(_swish(self.norm2(y))))
        return (self.nin_shortcut(x) if self.in_channels != self.out_channels else x) + y


class AttnBlock(nn.Module):
    """Single-head attention over the h*w positions of the feature map, softmax scale c**-0.5 (vq_model.py:316-350)."""

    def __init__(self, in_channels):
        super().__init__()
        self.norm = _gn(in_channels)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(in_channels, in_channels, 1, 1, 0) for _ in range(4))

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.norm(x)
        q, k, v = (m(y).reshape(b, 1, c, h * w).transpose(2, 3) for m in (self.q, self.k, self.v))   # [b, 1, hw, c]
        o = F.scaled_dot_product_attention(q, k, v, scale=float(int(c) ** (-0.5)))
        return x + self.proj_out(o.transpose(2, 3).reshape(b, c, h, w))


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv=True):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)

    def forward(self, x):
        if not self.with_conv:
            return F.avg_pool2d(x, kernel_size=2, stride=2)
        return self.conv(F.pad(x, (0, 1, 0, 1)))          # asymmetric zero pad: right / bottom only


def _level(block_in, block_out, n_res, with_attn, dropout):
    lvl = nn.Module()
    lvl.res, lvl.attn = nn.ModuleList(), nn.ModuleList()
    for _ in range(n_res):
        lvl.res.append(ResnetBlock(block_in, block_out, dropout))
        block_in = block_out
        if with_attn:
            lvl.attn.append(AttnBlock(block_in))
    return lvl


def _mid(ch, dropout):
    return nn.ModuleList([ResnetBlock(ch, ch, dropout), AttnBlock(ch), ResnetBlock(ch, ch, dropout)])


def _run_level(lvl, h):
    for i, res in enumerate(lvl.res):
        h = res(h)
        if len(lvl.attn) > 0:
            h = lvl.attn[i](h)
    return h


class Encoder(nn.Module):
    def __init__(self, in_channels=3, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, dropout=0.0, z_channels=256):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        widths = [ch] + [ch * m for m in ch_mult]
        self.conv_blocks = nn.ModuleList()
        for i in range(self.num_resolutions):
            last = i == self.num_resolutions - 1
            lvl = _level(widths[i], widths[i + 1], num_res_blocks, last, dropout)
            if not last:
                lvl.downsample = Downsample(widths[i + 1])
            self.conv_blocks.append(lvl)
        self.mid = _mid(widths[-1], dropout)
        self.norm_out = _gn(widths[-1])
        self.conv_out = nn.Conv2d(widths[-1], z_channels, 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for i, lvl in enumerate(self.conv_blocks):
            h = _run_level(lvl, h)
            if i != self.num_resolutions - 1:
                h = lvl.downsample(h)
        for m in self.mid:
            h = m(h)
        return self.conv_out(_swish(self.norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, z_channels=256, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, dropout=0.0, out_channels=3):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = _mid(block_in, dropout)
        self.conv_blocks = nn.ModuleList()
        for i in reversed(range(self.num_resolutions)):
            lvl = _level(block_in, ch * ch_mult[i], num_res_blocks + 1, i == self.num_resolutions - 1, dropout)
            block_in = ch * ch_mult[i]
            if i != 0:
                lvl.upsample = Upsample(block_in)
            self.conv_blocks.append(lvl)
        self.norm_out = _gn(block_in)
        self.conv_out = nn.Conv2d(block_in, out_channels, 3, 1, 1)

    @property
    def last_layer(self):
        return self.conv_out.weight

    def forward(self, z):
        h = self.conv_in(z)
        for m in self.mid:
            h = m(h)
        for i, lvl in enumerate(self.conv_blocks):
            h = _run_level(lvl, h)
            if i != self.num_resolutions - 1:
                h = lvl.upsample(h)
        return self.conv_out(_swish(self.norm_out(h)))


class VQModel(nn.Module):
    """vq_model.py:28-62.  ``encode`` ends in the vlg index search, ``decode_code`` starts with the fused
    gather + post_quant_conv kernel; both need a B200 (no CPU fallback).  ``decode`` / ``forward`` take any device the
    library convolutions run on."""

    def __init__(self, config: ModelArgs):
        super().__init__()
        self.config = config
        self.encoder = Encoder(ch_mult=config.encoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
        self.decoder = Decoder(ch_mult=config.decoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
        self.quantize = VectorQuantizer(config.codebook_size, config.codebook_embed_dim, config.commit_loss_beta,
                                        config.entropy_loss_ratio, config.codebook_l2_norm, config.codebook_show_usage)
        self.quant_conv = nn.Conv2d(config.z_channels, config.codebook_embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(config.codebook_embed_dim, config.z_channels, 1)

    def encode(self, x):
        h = self.quant_conv(self.encoder(x))
        return self.quantize(h.float().contiguous())

    def decode(self, quant):
        return self.decoder(self.post_quant_conv(quant))

    def decode_code(self, code_b, shape=None, channel_first=True):
        if shape is not None and channel_first and code_b.is_cuda:
            feat = self.quantize.decode_front(code_b, shape, self.post_quant_conv.weight, self.post_quant_conv.bias)
            return self.decoder(feat.to(self.decoder.conv_in.weight.dtype))
        quant_b = self.quantize.get_codebook_entry(code_b, shape, channel_first)
        return self.decode(quant_b)

    def forward(self, input):
        quant, diff, _ = self.encode(input)
        return self.decode(quant), diff


def VQ_8(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 2, 2, 4], decoder_ch_mult=[1, 2, 2, 4], **kwargs))


def VQ_16(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 1, 2, 2, 4], decoder_ch_mult=[1, 1, 2, 2, 4], **kwargs))


VQ_models = {'VQ-16': VQ_16, 'VQ-8': VQ_8}
