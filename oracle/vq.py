"""TEST INFRASTRUCTURE ONLY -- torch restatement of the reference's VQ index search
(tokenizer/tokenizer_image/vq_model.py ``VectorQuantizer``).  See oracle/__init__.py.
"""
import torch
import torch.nn.functional as F


def default_codebook(n_e=16384, e_dim=8, seed=0, l2_norm=True):
    """vq_model.py:207-210: uniform(-1/n_e, 1/n_e) then l2-normalise."""
    g = torch.Generator().manual_seed(seed)
    cb = (torch.rand(n_e, e_dim, generator=g) * 2 - 1) / n_e
    return F.normalize(cb, p=2, dim=-1) if l2_norm else cb


@torch.no_grad()
def quantize(z, codebook, l2_norm=True, chunk=65536):
    """vq_model.py:215-259 in eval mode.

    z [B,C,H,W] fp32 -> (z_q [B,C,H,W] fp32, indices int64 [B*H*W]).
    d = |z|^2 + |e|^2 - 2 z.e^T is materialised per chunk of latents exactly as the
    reference does per call (the un-chunked 1M-latent call needs a 68 GB matrix;
    SURVEY.md 8c item 5 shows chunking does not change any index)."""
    zl = z.permute(0, 2, 3, 1).contiguous()
    flat = zl.view(-1, codebook.shape[1])
    if l2_norm:
        zl = F.normalize(zl, p=2, dim=-1)
        flat = F.normalize(flat, p=2, dim=-1)
        emb = F.normalize(codebook, p=2, dim=-1)
    else:
        emb = codebook
    idx = torch.empty(flat.shape[0], dtype=torch.int64, device=z.device)
    e2 = torch.sum(emb ** 2, dim=1)
    for s in range(0, flat.shape[0], chunk):
        f = flat[s:s + chunk]
        d = torch.sum(f ** 2, dim=1, keepdim=True) + e2 - 2 * torch.einsum("bd,dn->bn", f, emb.t())
        idx[s:s + chunk] = torch.argmin(d, dim=1)
    z_q = emb[idx].view(zl.shape)
    z_q = zl + (z_q - zl)                      # straight-through form, vq_model.py:254
    return z_q.permute(0, 3, 1, 2), idx


@torch.no_grad()
def codebook_entry(indices, codebook, shape=None, channel_first=True, l2_norm=True):
    """vq_model.py:261-276."""
    emb = F.normalize(codebook, p=2, dim=-1) if l2_norm else codebook
    z_q = emb[indices]
    if shape is not None:
        if channel_first:
            z_q = z_q.reshape(shape[0], shape[2], shape[3], shape[1]).permute(0, 3, 1, 2).contiguous()
        else:
            z_q = z_q.view(shape)
    return z_q
