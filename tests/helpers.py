"""Shared helpers for the tests: fixture -> (Geometry, weights) decoding."""
import numpy as np
import torch

from oracle import ar_decode as O


def weights_from(npz, dtype=torch.float32, device="cpu"):
    return {k[2:]: torch.from_numpy(np.asarray(npz[k])).to(device=device, dtype=dtype)
            for k in npz.files if k.startswith("w.")}


TINY_C2I = O.Geometry(dim=128, n_layer=2, n_head=2, vocab_size=512, block_size=16, multiple_of=64,
                      cls_token_num=1, num_classes=10, model_type="c2i")
TINY_T2I = O.Geometry(dim=128, n_layer=2, n_head=2, vocab_size=256, block_size=16, multiple_of=64,
                      cls_token_num=120, caption_dim=32, model_type="t2i")
TINY_T2V = O.Geometry(dim=128, n_layer=2, n_head=2, vocab_size=64, block_size=16, multiple_of=64,
                      cls_token_num=120, caption_dim=64, model_type="t2v", vae_embed_dim=8, num_frames=5,
                      t_downsample_size=4)
