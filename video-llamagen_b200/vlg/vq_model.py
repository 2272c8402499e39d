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
        scratch = torch.empty(self.n_e * (Cc + 1), dtype=torch.float32, device=z.device)
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
