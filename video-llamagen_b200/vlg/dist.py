"""Sample-sharded multi-GPU decode: one process per GPU, no communication inside the decode loop, one all_gather of
the int32 code indices at the end.  Independent samples never interact -- each row owns its KV cache and
a CFG pair (b, b+B) stays on one GPU -- so this is the whole multi-GPU story of the path; the reference's missing
``sample_c2i_ddp.py`` followed the same pattern (rank-strided work, gather at the end:
tokenizer/tokenizer_image/reconstruction_vq_ddp.py:49-178)."""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def rank_seed(global_seed, rank=None, world_size=None):
    """Per-rank RNG seed, the reference's convention (autoregressive/train/train_t2i.py:70)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return global_seed * world_size + rank


def shard_bounds(n, rank, world_size):
    """Contiguous [lo, hi) slice of n samples for `rank`; ragged tails go to the lowest ranks."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(cond, emb_masks=None):
    rank, w = world()
    lo, hi = shard_bounds(cond.shape[0], rank, w)
    return cond[lo:hi], (None if emb_masks is None else emb_masks[lo:hi])


def gather_indices(local_tokens, n_total):
    """all_gather of int32 [B_local, N] code indices -> [n_total, N] on every rank (ragged shards are padded)."""
    rank, w = world()
    if w == 1:
        return local_tokens
    per = (n_total + w - 1) // w
    N = local_tokens.shape[1]
    padded = torch.zeros(per, N, dtype=local_tokens.dtype, device=local_tokens.device)
    padded[: local_tokens.shape[0]] = local_tokens
    out = torch.empty(w * per, N, dtype=local_tokens.dtype, device=local_tokens.device)
    if local_tokens.is_cuda:
        dist.all_gather_into_tensor(out, padded)
    else:                                                   # gloo (CPU tests)
        parts = [torch.empty_like(padded) for _ in range(w)]
        dist.all_gather(parts, padded)
        out = torch.cat(parts)
    rows = []
    for r in range(w):
        lo, hi = shard_bounds(n_total, r, w)
        rows.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(rows)


def generate_sharded(model, cond, max_new_tokens, emb_masks=None, generate_fn=None, global_seed=0, **kw):
    """generate() over a global batch: each rank decodes its shard, every rank returns all [B, N] indices."""
    if generate_fn is None:
        from .generate import generate as generate_fn
    local_cond, local_masks = shard_batch(cond, emb_masks)
    if local_cond.shape[0] > 0:
        toks = generate_fn(model, local_cond, max_new_tokens, emb_masks=local_masks, seed=rank_seed(global_seed), **kw)
    else:
        toks = torch.zeros(0, max_new_tokens, dtype=torch.int32, device=cond.device)
    return gather_indices(toks, cond.shape[0])
