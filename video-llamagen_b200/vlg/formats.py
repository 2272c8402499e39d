"""On-disk formats on either side of the decode path, as the reference's trainers, samplers and evaluator read and write them
(SURVEY.md 8f rank 4).  numpy only: nothing here touches the GPU; the arrays go to / come from ``vlg.generate`` and
``VQModel.encode`` / ``decode_code``.

* image code files      ``<code_dir>/<name>.npy``  int64 ``[1, h*w]``; flip / crop augmented: ``[1, n_aug, h*w]``
                        (written by autoregressive/train/extract_codes_t2i.py:115-122 and extract_codes_c2i.py, read by
                        dataset/imagenet.py:45-50 and tools/check_image_codes.py:28-34)
* label files           ``<label_dir>/<i>.npy``    int64 ``[1]``                                  (dataset/imagenet.py:49)
* T5 caption features   ``<t5_dir>/<name>.npy``    float ``[1, len, 2048]``; the loader LEFT-pads to ``max_len`` tokens and
                        marks the valid tail in ``emb_mask``                                      (dataset/t2i.py:119-141)
* FID sample archive    ``<sample_dir>.npz``       ``arr_0`` uint8 ``[N, H, W, 3]``
                        (tokenizer/tokenizer_image/reconstruction_vq_ddp.py:24-39; read by evaluations/c2i/evaluator.py)
"""
import os
from typing import Iterable, Optional, Tuple

import numpy as np


# ----------------------------------------------------------------------------------------------- image codes
def save_codes(path: str, codes) -> None:
    """codes: int array ``[h*w]``, ``[1, h*w]`` (one image) or ``[n_aug, h*w]`` / ``[1, n_aug, h*w]`` (augmented views of one
    image) -> the ``[1, ...]`` int64 layout of extract_codes_t2i.py:118-122 (batch size is 1 there: "important!")."""
    x = np.asarray(codes)
    if not np.issubdtype(x.dtype, np.integer):
        raise TypeError(f"codes must be integers, got {x.dtype}")
    if x.ndim == 1:
        x = x[None]
    elif x.ndim == 2 and x.shape[0] != 1:
        x = x[None]
    if x.ndim not in (2, 3) or x.shape[0] != 1:
        raise ValueError(f"codes of one image are [1, h*w] or [1, n_aug, h*w], got {x.shape}")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    np.save(path, x.astype(np.int64))


def load_codes(path: str, aug_idx: Optional[int] = None) -> np.ndarray:
    """-> int64 ``[1, h*w]``.  An augmented file ``[1, n_aug, h*w]`` needs ``aug_idx`` (dataset/imagenet.py:46-48 draws it at
    random); a plain file ignores it."""
    x = np.load(path)
    if x.ndim == 3:
        if aug_idx is None:
            raise ValueError(f"{path} holds {x.shape[1]} augmented views: pass aug_idx")
        x = x[:, aug_idx]
    if x.ndim != 2 or x.shape[0] != 1:
        raise ValueError(f"{path}: expected [1, h*w] or [1, n_aug, h*w], got {x.shape}")
    return x.astype(np.int64, copy=False)


def codes_for_decode(path: str, latent_dim: int, latent_size: int) -> Tuple[np.ndarray, Tuple[int, int, int, int]]:
    """tools/check_image_codes.py:28-34: -> (flat indices ``[n*h*w]``, ``qzshape`` for ``VQModel.decode_code``): every augmented
    view of the file becomes one image of the decoded batch."""
    x = np.load(path)
    n = x.shape[1] if x.ndim == 3 else 1
    if x.reshape(-1).shape[0] != n * latent_size * latent_size:
        raise ValueError(f"{path}: {x.shape} does not hold {n} x {latent_size}x{latent_size} codes")
    return x.reshape(-1).astype(np.int64, copy=False), (n, latent_dim, latent_size, latent_size)


def save_label(path: str, label: int) -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    np.save(path, np.asarray([label], dtype=np.int64))


# ----------------------------------------------------------------------------------------------- T5 features
def save_t5_feature(path: str, feat) -> None:
    """feat ``[len, dim]`` or ``[1, len, dim]`` -> ``[1, len, dim]`` float32 (language/extract_t5_feature.py layout)."""
    x = np.asarray(feat, dtype=np.float32)
    if x.ndim == 2:
        x = x[None]
    if x.ndim != 3 or x.shape[0] != 1:
        raise ValueError(f"a caption feature is [1, len, dim], got {x.shape}")
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    np.save(path, x)


def pad_t5_feature(feat, max_len: int = 120, dim: int = 2048) -> Tuple[np.ndarray, np.ndarray]:
    """dataset/t2i.py:119-133: the first ``min(max_len, len)`` tokens go to the END of a zero ``[1, max_len, dim]`` buffer and
    ``emb_mask [max_len]`` is 1 on that tail -- the layout ``generate(..., emb_masks=...)`` and the masked prefix attention
    expect (sample_t2i.py:103-118 builds the same thing from the live T5)."""
    x = np.asarray(feat, dtype=np.float32)
    if x.ndim == 2:
        x = x[None]
    if x.ndim != 3 or x.shape[0] != 1 or x.shape[2] != dim:
        raise ValueError(f"expected [1, len, {dim}], got {x.shape}")
    n = min(max_len, x.shape[1])
    out = np.zeros((1, max_len, dim), dtype=np.float32)
    mask = np.zeros((max_len,), dtype=np.float32)
    if n > 0:
        out[:, -n:] = x[:, :n]
        mask[-n:] = 1
    return out, mask


def load_t5_feature(path: str, max_len: int = 120, dim: int = 2048) -> Tuple[np.ndarray, np.ndarray]:
    return pad_t5_feature(np.load(path), max_len, dim)


def training_attn_mask(emb_mask, max_seq_length: int) -> np.ndarray:
    """dataset/t2i.py:134-140: causal mask ``[1, S, S]`` bool whose first ``T`` columns are gated by ``emb_mask`` and whose
    diagonal is always on (a fully padded row still attends to itself).  The decode kernels evaluate exactly this predicate
    per (query, key) instead of reading the matrix."""
    m = np.asarray(emb_mask, dtype=np.float32)
    T, S = m.shape[0], max_seq_length
    a = np.tril(np.ones((S, S), dtype=np.float32))
    a[:, :T] = a[:, :T] * m[None, :]
    eye = np.eye(S, dtype=np.float32)
    a = a * (1 - eye) + eye
    return a[None].astype(bool)


# ----------------------------------------------------------------------------------------------- FID sample archive
def images_to_uint8(samples) -> np.ndarray:
    """decoder output ``[B, 3, H, W]`` in [-1, 1] (torch tensor or array) -> uint8 ``[B, H, W, 3]``, the conversion of
    reconstruction_vq_ddp.py:137: ``clamp(127.5 * x + 128.0, 0, 255)`` then truncation to uint8."""
    if hasattr(samples, "detach"):
        samples = samples.detach().float().cpu().numpy()
    x = np.asarray(samples, dtype=np.float32)
    if x.ndim != 4 or x.shape[1] != 3:
        raise ValueError(f"expected [B, 3, H, W], got {x.shape}")
    x = np.clip(127.5 * x + 128.0, 0, 255)
    return np.ascontiguousarray(x.transpose(0, 2, 3, 1)).astype(np.uint8)


class NpzSampleWriter:
    """Collects generated images and writes the archive the reference's FID evaluator reads: ``arr_0`` uint8 ``[N, H, W, 3]``
    (create_npz_from_sample_folder, reconstruction_vq_ddp.py:24-39, without the round trip through N png files).  ``num`` caps
    the archive (50 000 for the ImageNet protocol, GETTING_STARTED.md:60): samples past it are dropped, as the reference
    drops the padding of the last DDP batch."""

    def __init__(self, num: int = 50_000):
        self.num = int(num)
        self._chunks = []
        self._count = 0
        self._hw = None

    def __len__(self) -> int:
        return self._count

    def add(self, images) -> int:
        """images: uint8 ``[B, H, W, 3]`` or decoder output ``[B, 3, H, W]`` in [-1, 1]; returns how many were kept."""
        x = images
        if hasattr(x, "detach") or np.asarray(x).dtype != np.uint8:
            x = images_to_uint8(x)
        x = np.asarray(x)
        if x.ndim != 4 or x.shape[3] != 3:
            raise ValueError(f"expected uint8 [B, H, W, 3], got {x.shape}")
        if self._hw is None:
            self._hw = x.shape[1:3]
        elif tuple(x.shape[1:3]) != tuple(self._hw):
            raise ValueError(f"image size changed from {self._hw} to {x.shape[1:3]}")
        keep = max(0, min(x.shape[0], self.num - self._count))
        if keep:
            self._chunks.append(np.ascontiguousarray(x[:keep]))
            self._count += keep
        return keep

    def write(self, npz_path: str, require_full: bool = False) -> str:
        if require_full and self._count != self.num:
            raise ValueError(f"{self._count} samples collected, {self.num} required")
        if not self._chunks:
            raise ValueError("no samples")
        samples = np.concatenate(self._chunks, axis=0)
        assert samples.shape == (self._count, samples.shape[1], samples.shape[2], 3)
        d = os.path.dirname(npz_path)
        if d:
            os.makedirs(d, exist_ok=True)
        np.savez(npz_path, arr_0=samples)
        return npz_path


def gather_rank_samples(per_rank: Iterable[np.ndarray], total: int) -> np.ndarray:
    """Interleave what the DDP samplers produce: rank r generated global samples r, r + W, r + 2W, ... (``index = i *
    world_size + rank + total``, reconstruction_vq_ddp.py:141); -> the first ``total`` samples in global order."""
    per_rank = [np.asarray(a) for a in per_rank]
    W = len(per_rank)
    n = min(a.shape[0] for a in per_rank)
    out = np.stack(per_rank, axis=1).reshape((n * W,) + per_rank[0].shape[1:]) if all(a.shape[0] == n for a in per_rank) else None
    if out is None:
        raise ValueError("every rank must hold the same number of samples")
    return out[:total]
