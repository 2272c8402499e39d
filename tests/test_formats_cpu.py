"""On-disk formats around the decode path (vlg/formats.py) against the reference's readers / writers
(dataset/t2i.py:119-141, dataset/imagenet.py:45-50, tools/check_image_codes.py:28-34, reconstruction_vq_ddp.py:24-39,137)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "video-llamagen_b200")]
from vlg import formats as F  # noqa: E402


def test_code_files_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    one = rng.integers(0, 16384, size=(576,))
    p = str(tmp_path / "a" / "7.npy")
    F.save_codes(p, one)
    raw = np.load(p)
    assert raw.shape == (1, 576) and raw.dtype == np.int64           # extract_codes_t2i.py:118-122
    assert np.array_equal(F.load_codes(p)[0], one)
    aug = rng.integers(0, 16384, size=(10, 256))
    p2 = str(tmp_path / "b.npy")
    F.save_codes(p2, aug)
    assert np.load(p2).shape == (1, 10, 256)                         # flip / ten-crop layout, dataset/imagenet.py:46-48
    with pytest.raises(ValueError):
        F.load_codes(p2)
    assert np.array_equal(F.load_codes(p2, aug_idx=3)[0], aug[3])
    # what the reference's dataset does with the same file
    feats = np.load(p2)
    assert np.array_equal(torch.from_numpy(feats[:, 3]).numpy(), F.load_codes(p2, 3))
    idx, qz = F.codes_for_decode(p2, latent_dim=8, latent_size=16)    # check_image_codes.py:28-34
    assert qz == (10, 8, 16, 16) and idx.shape == (2560,) and np.array_equal(idx, aug.reshape(-1))
    idx, qz = F.codes_for_decode(p, 8, 24)
    assert qz == (1, 8, 24, 24)
    with pytest.raises(TypeError):
        F.save_codes(p, one.astype(np.float32))
    with pytest.raises(ValueError):
        F.codes_for_decode(p, 8, 16)


def _reference_t5_padding(t5_feat, max_len, dim, max_seq_length):
    """dataset/t2i.py:119-141 verbatim semantics, on torch tensors as the reference does it."""
    t5_feat_padding = torch.zeros((1, max_len, dim))
    t5_feat = torch.from_numpy(t5_feat)
    feat_len = min(max_len, t5_feat.shape[1])
    t5_feat_padding[:, -feat_len:] = t5_feat[:, :feat_len]
    emb_mask = torch.zeros((max_len,))
    emb_mask[-feat_len:] = 1
    attn_mask = torch.tril(torch.ones(max_seq_length, max_seq_length))
    attn_mask[:, :max_len] = attn_mask[:, :max_len] * emb_mask.unsqueeze(0)
    eye = torch.eye(max_seq_length, max_seq_length)
    attn_mask = (attn_mask * (1 - eye) + eye).unsqueeze(0).to(torch.bool)
    return t5_feat_padding.numpy(), emb_mask.numpy(), attn_mask.numpy()


@pytest.mark.parametrize("length", [1, 37, 120, 200])
def test_t5_feature_padding_matches_the_reference_loader(tmp_path, length):
    rng = np.random.default_rng(length)
    feat = rng.standard_normal((1, length, 2048)).astype(np.float32)
    p = str(tmp_path / "t5" / "x.npy")
    F.save_t5_feature(p, feat[0])
    assert np.load(p).shape == (1, length, 2048)
    pad, mask = F.load_t5_feature(p, 120, 2048)
    ref_pad, ref_mask, ref_attn = _reference_t5_padding(feat, 120, 2048, 120 + 256)
    assert np.array_equal(pad, ref_pad) and np.array_equal(mask, ref_mask)
    assert np.array_equal(F.training_attn_mask(mask, 120 + 256), ref_attn)
    n = min(length, 120)
    assert mask[-n:].all() and not mask[:-n].any()                   # left padding: ones at the END (sample_t2i.py:103-118)


def test_npz_archive_is_what_the_evaluator_reads(tmp_path):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(5, 3, 8, 8, generator=g) * 2.4 - 1.2              # decoder output, a little outside [-1, 1]
    ref = torch.clamp(127.5 * x + 128.0, 0, 255).permute(0, 2, 3, 1).to("cpu", dtype=torch.uint8).numpy()   # reconstruction_vq_ddp.py:137
    assert np.array_equal(F.images_to_uint8(x), ref)
    w = F.NpzSampleWriter(num=8)
    assert w.add(x) == 5
    assert w.add(ref) == 3 and len(w) == 8                           # capped at num: the tail of the last batch is dropped
    assert w.add(ref) == 0
    path = w.write(str(tmp_path / "samples" / "run.npz"), require_full=True)
    arr = np.load(path)["arr_0"]
    assert arr.dtype == np.uint8 and arr.shape == (8, 8, 8, 3)
    assert np.array_equal(arr[:5], ref) and np.array_equal(arr[5:], ref[:3])
    with pytest.raises(ValueError):
        F.NpzSampleWriter(4).write(str(tmp_path / "empty.npz"))
    w2 = F.NpzSampleWriter(4)
    w2.add(ref[:2])
    with pytest.raises(ValueError):
        w2.write(str(tmp_path / "short.npz"), require_full=True)
    with pytest.raises(ValueError):
        w2.add(np.zeros((1, 4, 4, 3), np.uint8))                     # size change


def test_rank_interleave_matches_the_ddp_sample_index():
    W, n_batches, n = 4, 3, 5
    per_rank = [[] for _ in range(W)]
    expect = {}
    total = 0
    for _ in range(n_batches):                                       # reconstruction_vq_ddp.py:139-150
        for rank in range(W):
            for i in range(n):
                index = i * W + rank + total
                tag = np.array([rank, len(per_rank[rank])])
                per_rank[rank].append(tag)
                expect[index] = tag
        total += n * W
    out = F.gather_rank_samples([np.stack(a) for a in per_rank], total=50)
    assert out.shape == (50, 2)
    for k in range(50):
        assert np.array_equal(out[k], expect[k])
    with pytest.raises(ValueError):
        F.gather_rank_samples([np.zeros((3, 1)), np.zeros((2, 1))], 4)
