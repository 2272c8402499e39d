"""Drop-in for the reference's ``autoregressive/models/generate.py``.

``generate()`` runs the whole loop inside libvlg.so (prefill, N-1 graph-replayed decode steps, fused CFG +
top-k/top-p + sampling, device-resident loop state).  The finer-grained functions keep the reference's
signatures for callers that drive the loop themselves; they compose the same kernels from Python.

Sampling randomness: ``torch.multinomial``'s single-draw path is ``argmax(probs / Exp(1))``; the kernel does the
same with Exp(1) draws from a Philox stream seeded from torch's default CUDA generator (``torch.manual_seed``
controls it), so runs are reproducible but not draw-for-draw identical to torch -- pass ``noise=`` to
:func:`sample` for draw-for-draw comparisons.
"""
import ctypes as C

import torch

from . import _lib as L


def _seed(device):
    """A fresh 63-bit seed drawn from torch's CPU generator (so torch.manual_seed makes runs reproducible)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def _run_sampler(logits2d, rows, B, cfg_scale, guided, temperature, top_k, top_p, sample_logits, noise=None,
                 want_probs=True, seed=None, offset=0):
    """One launch of the fused sampler.  ``want_probs=False`` with ``noise=None`` and ``sample_logits=True`` selects the
    kernel branch the in-library decode loop runs (no probability write, Philox draws, log-domain race); ``offset`` is
    the Philox stream offset the loop derives from the step index."""
    L.require_cuda(logits2d, "sample")
    V = logits2d.shape[-1]
    p = L.SampleParams(float(cfg_scale), int(guided), float(temperature), int(top_k or 0), float(top_p),
                       int(bool(sample_logits)), int(seed if seed is not None else _seed(logits2d.device)), int(offset))
    idx = torch.empty(B, dtype=torch.int64, device=logits2d.device)
    probs = torch.empty(B, V, dtype=torch.float32, device=logits2d.device) if want_probs else None
    if noise is not None:
        noise = noise.to(device=logits2d.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(logits2d.device):
        L.check(L.lib().vlg_cfg_sample(L.ptr(logits2d), rows, B, V, C.byref(p), L.ptr(noise), L.ptr(idx), L.ptr(probs),
                                       L.stream()))
    return idx.view(B, 1), probs


def top_k_top_p_filtering(logits, top_k: int = 0, top_p: float = 1.0, filter_value: float = -float("Inf"),
                          min_tokens_to_keep: int = 1):
    """generate.py:16-54, in place on logits [B, V] (fp32, CUDA)."""
    if filter_value != -float("Inf") or min_tokens_to_keep != 1:
        raise NotImplementedError("only the reference's defaults (filter_value=-inf, min_tokens_to_keep=1) are used on the path")
    L.require_cuda(logits, "top_k_top_p_filtering")
    assert logits.dim() == 2 and logits.dtype == torch.float32
    out = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        L.check(L.lib().vlg_top_k_top_p(L.ptr(logits), L.ptr(out), logits.shape[0], logits.shape[1], int(top_k or 0),
                                        float(top_p), L.stream()))
    logits.copy_(out)
    return logits


def sample(logits, temperature: float = 1.0, top_k: int = 0, top_p: float = 1.0, sample_logits=True, noise=None):
    """generate.py:57-66: logits [B, S, V] -> (idx int64 [B, 1], probs fp32 [B, V])."""
    last = logits[:, -1, :].to(torch.float32).contiguous()
    return _run_sampler(last, last.shape[0], last.shape[0], 1.0, 0, temperature, top_k, top_p, sample_logits, noise)


def logits_to_probs(logits, temperature: float = 1.0, top_p: float = 1.0, top_k: int = None, **kwargs):
    """generate.py:69-74 for logits [B, V]."""
    l2 = logits.to(torch.float32).contiguous()
    return _run_sampler(l2, l2.shape[0], l2.shape[0], 1.0, 0, temperature, top_k, top_p, False)[1]


def _guided_sample(logits, cfg_scale, cfg_flag, noise=None, **sampling_kwargs):
    """CFG mix (generate.py:81-82 / 95-99) fused with sample() on the last position of logits [R, S, V]."""
    last = logits[:, -1, :].to(torch.float32).contiguous()
    R = last.shape[0]
    if cfg_scale > 1.0:
        return _run_sampler(last, R, R // 2, cfg_scale, 1 if cfg_flag else 0, sampling_kwargs.get("temperature", 1.0),
                            sampling_kwargs.get("top_k", 0), sampling_kwargs.get("top_p", 1.0),
                            sampling_kwargs.get("sample_logits", True), noise)
    return _run_sampler(last, R, R, 1.0, 0, sampling_kwargs.get("temperature", 1.0), sampling_kwargs.get("top_k", 0),
                        sampling_kwargs.get("top_p", 1.0), sampling_kwargs.get("sample_logits", True), noise)


def prefill(model, cond_idx: torch.Tensor, input_pos: torch.Tensor, cfg_scale: float, **sampling_kwargs):
    """generate.py:77-86."""
    logits, _ = model(None, cond_idx, input_pos)
    return _guided_sample(logits, cfg_scale, True, **sampling_kwargs)[0]


def decode_one_token(model, x: torch.Tensor, input_pos: torch.Tensor, cfg_scale: float, cfg_flag: bool,
                     **sampling_kwargs):
    """generate.py:89-102."""
    assert input_pos.shape[-1] == 1
    if cfg_scale > 1.0:
        x = torch.cat([x, x])
    logits, _ = model(x, cond_idx=None, input_pos=input_pos)
    return _guided_sample(logits, cfg_scale, cfg_flag, **sampling_kwargs)


def decode_n_tokens(model, cur_token: torch.Tensor, input_pos: torch.Tensor, num_new_tokens: int, cfg_scale: float,
                    cfg_interval: int, **sampling_kwargs):
    """generate.py:105-124 (Python-driven loop; generate() uses the in-library graph loop instead)."""
    new_tokens, new_probs = [], []
    cfg_flag = True
    for i in range(num_new_tokens):
        if cfg_interval > -1 and i > cfg_interval:
            cfg_flag = False
        next_token, next_prob = decode_one_token(model, cur_token, input_pos, cfg_scale, cfg_flag, **sampling_kwargs)
        input_pos += 1
        new_tokens.append(next_token.clone())
        new_probs.append(next_prob.clone())
        cur_token = next_token.view(-1, 1)
    return new_tokens, new_probs


@torch.no_grad()
def generate(model, cond, max_new_tokens, emb_masks=None, cfg_scale=1.0, cfg_interval=-1, **sampling_kwargs):
    """generate.py:127-180 -> int32 [B, max_new_tokens] on cond's device."""
    if model.model_type == 'c2i':
        T = 1
        cond_in = cond.to(torch.int64).contiguous()
    elif model.model_type == 't2i':
        T = cond.shape[1]
        assert T == model.cls_token_num
        cond_in = cond.to(model.tok_embeddings.weight.dtype).contiguous()
    else:
        raise Exception("please check model type")
    L.require_cuda(cond_in, "generate")
    B = cond.shape[0]
    rows = 2 * B if cfg_scale > 1.0 else B
    if emb_masks is not None:
        assert emb_masks.shape[0] == B and emb_masks.shape[-1] == T
    eng = model.engine(rows, T + max_new_tokens)
    model.max_batch_size, model.max_seq_length = rows, (T + max_new_tokens + 7) // 8 * 8
    return eng.generate(cond_in, max_new_tokens, emb_masks, cfg_scale, cfg_interval,
                        sampling_kwargs.get("temperature", 1.0), sampling_kwargs.get("top_k", 0),
                        sampling_kwargs.get("top_p", 1.0), sampling_kwargs.get("sample_logits", True),
                        sampling_kwargs.get("seed", None) if sampling_kwargs.get("seed", None) is not None
                        else _seed(cond.device))
