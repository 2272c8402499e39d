"""GPU: per-kernel parity through the C ABI against the oracle (torch ops on the same device / CPU)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import ar_decode as O
from oracle import vq as OV
from vlg import _lib as L
from vlg.vq_model import VectorQuantizer
from vlg import generate as G

pytestmark = pytest.mark.gpu
DEV = "cuda"


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


# ------------------------------------------------------------------------------------------ VQ argmin
def test_vq_fixture_indices_bit_exact(golden):
    f = golden("vq.npz")
    vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval()
    vq.embedding.weight.data = torch.from_numpy(f["codebook"])
    vq = vq.to(DEV)
    z_q, _, (_, _, idx) = vq(torch.from_numpy(f["z"]).to(DEV))
    assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), f["idx"])
    assert np.abs(z_q.cpu().numpy() - f["z_q"]).max() < 1e-6
    e = vq.get_codebook_entry(idx, (4, 8, 16, 16), channel_first=True)
    assert np.abs(e.cpu().numpy() - f["entry"]).max() < 1e-6
    e2 = vq.get_codebook_entry(idx, None)
    assert e2.shape == (1024, 8)
    vq2 = VectorQuantizer(1000, 8, 0.25, 0.0, False, False).eval()
    vq2.embedding.weight.data = torch.from_numpy(f["codebook2"])
    vq2 = vq2.to(DEV)
    _, _, (_, _, idx2) = vq2(torch.from_numpy(f["z2"]).to(DEV))
    assert np.array_equal(idx2.cpu().numpy(), f["idx2"])                # ragged HW = 35, no l2 norm


def test_vq_64k_latents_vs_oracle_and_edge_cases():
    cb = OV.default_codebook(16384, 8, seed=0)
    g = torch.Generator().manual_seed(1234)
    z = torch.randn(64, 8, 32, 32, generator=g)
    vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval()
    vq.embedding.weight.data = cb.clone()
    vq = vq.to(DEV)
    _, _, (_, _, idx) = vq(z.to(DEV))
    _, ref = OV.quantize(z, cb)
    mism = (idx.cpu() != ref).nonzero().flatten()
    assert mism.numel() == 0, f"{mism.numel()} index mismatches, first at {mism[:5].tolist()}"
    # empty batch and a single latent
    _, _, (_, _, e) = vq(torch.zeros(0, 8, 4, 4, device=DEV))
    assert e.numel() == 0
    _, _, (_, _, one) = vq(z[:1, :, :1, :1].contiguous().to(DEV))
    assert one.item() == ref[0].item()
    # duplicated codebook rows: the lowest index must win (torch.argmin tie rule)
    vq.embedding.weight.data[9000] = vq.embedding.weight.data[17]
    _, _, (_, _, idx_d) = vq(z[:2].to(DEV))
    assert not (idx_d == 9000).any()


def test_vq_one_million_latents_properties():
    """cfg-5 size (1024x8x32x32): size-independent properties instead of a CPU replay."""
    cb = OV.default_codebook(16384, 8, seed=0)
    vq = VectorQuantizer(16384, 8, 0.25, 0.0, True, True).eval()
    vq.embedding.weight.data = cb.clone()
    vq = vq.to(DEV)
    z = torch.randn(1024, 8, 32, 32, generator=torch.Generator(device=DEV).manual_seed(5), device=DEV)
    z_q, _, (_, _, idx) = vq(z)
    assert idx.shape == (1 << 20,) and int(idx.min()) >= 0 and int(idx.max()) < 16384
    # index for index against the reference algorithm (the [chunk, 16384] distance matrix + argmin, 16 chunks of 65 536
    # latents) run on this GPU.  The library's fp32 GEMM sums in another order than the kernel, so a disagreement is
    # legitimate only where the two best codes tie to within fp32 rounding of d -- verified in fp64, and counted.
    _, ref = OV.quantize(z, cb.to(DEV), chunk=65536)
    mism = (idx != ref).nonzero().flatten()
    if mism.numel():
        zn64 = torch.nn.functional.normalize(z.permute(0, 2, 3, 1).reshape(-1, 8)[mism].double(), dim=-1)
        en64 = torch.nn.functional.normalize(cb.to(DEV).double(), dim=-1)
        gap = ((zn64 * en64[idx[mism]]).sum(-1) - (zn64 * en64[ref[mism]]).sum(-1)).abs() * 2
        assert mism.numel() <= 4 and float(gap.max()) < 2.5e-7, (mism.numel(), float(gap.max()), mism[:5].tolist())
    # idempotence: quantising the quantised latents returns the same codes
    _, _, (_, _, idx2) = vq(z_q)
    assert torch.equal(idx, idx2)
    # the chosen code maximises the cosine similarity (within fp32 noise) on a 4096-latent sample
    sel = torch.randperm(1 << 20, generator=torch.Generator().manual_seed(0))[:4096].to(DEV)
    zn = torch.nn.functional.normalize(z.permute(0, 2, 3, 1).reshape(-1, 8)[sel].double(), dim=-1)
    sim = zn @ torch.nn.functional.normalize(cb.to(DEV).double(), dim=-1).t()
    best = sim.max(dim=1).values
    got = sim.gather(1, idx[sel].view(-1, 1)).flatten()
    assert float((best - got).max()) < 1e-6
    # permuting the latents permutes the indices
    perm = torch.randperm(1024, generator=torch.Generator().manual_seed(1)).to(DEV)
    _, _, (_, _, idx_p) = vq(z[perm].contiguous())
    assert torch.equal(idx_p.view(1024, -1), idx.view(1024, -1)[perm])


def test_vqmodel_encode_and_decode_code_vs_reference_fixture(golden):
    """The tokenizer surface the samplers call (vq_model.py:41-55): VQ_models['VQ-16'] rebuilt from the seed, encode() ->
    the index search kernel, decode_code() -> fused gather + post_quant_conv kernel -> decoder CNN (library convolutions,
    fp32, TF32 off).  Tolerances: the fused front is 8-term fp32 dot products (1e-5 absolute); the CNN outputs are compared
    with the reference's CPU fp32 run at 1e-3 of the output range (summation order of cuDNN vs MKL-DNN over ~30 layers)."""
    from vlg.vq_model import VQ_models
    f = golden("vqmodel.npz")
    torch.manual_seed(0)
    m = VQ_models["VQ-16"](codebook_size=16384, codebook_embed_dim=8).eval().to(DEV)
    x = torch.from_numpy(f["x"]).to(DEV)
    with torch.no_grad():
        h = m.quant_conv(m.encoder(x))
        quant, _, (_, _, idx) = m.encode(x)
    assert float((h.cpu() - torch.from_numpy(f["h"])).abs().max()) < 1e-3 * float(np.abs(f["h"]).max())
    _, ref_idx = OV.quantize(h.cpu(), m.quantize.embedding.weight.detach().cpu())
    assert torch.equal(idx.cpu(), ref_idx)                             # bit-exact on the same encoder output
    assert (idx.cpu().numpy() == f["idx"]).mean() >= 0.9               # the fixture's h differs in the last bits
    assert quant.shape == (2, 8, 4, 4)
    codes = torch.from_numpy(f["codes"]).to(DEV)
    feat = m.quantize.decode_front(codes, (2, 8, 8, 8), m.post_quant_conv.weight, m.post_quant_conv.bias)
    assert float((feat.cpu() - torch.from_numpy(f["feat"])).abs().max()) < 1e-5
    with torch.no_grad():
        img = m.decode_code(codes, (2, 8, 8, 8))
        img2 = m.decode(m.quantize.get_codebook_entry(codes, (2, 8, 8, 8)))
    rng = float(np.abs(f["img"]).max())
    assert img.shape == (2, 3, 128, 128)
    assert float((img.cpu() - torch.from_numpy(f["img"])).abs().max()) < 1e-3 * rng
    assert float((img - img2).abs().max()) < 1e-4 * rng               # fused front == gather followed by the 1x1 conv


# ------------------------------------------------------------------------------------------ RMSNorm
@pytest.mark.parametrize("dtype,rows,dim", [(torch.float32, 8, 768), (torch.bfloat16, 128, 1024),
                                            (torch.bfloat16, 5, 1536), (torch.float32, 3, 128)])
def test_rmsnorm(dtype, rows, dim):
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(rows, dim, generator=g) * 2).to(DEV, dtype)
    w = (1 + 0.1 * torch.randn(dim, generator=g)).to(DEV, dtype)
    out = torch.empty_like(x)
    L.check(L.lib().vlg_rmsnorm(L.ptr(x), L.ptr(w), L.ptr(out), rows, dim, 1e-5, L.dtype_code(dtype), L.stream()))
    ref = O.rms_norm(x, w, 1e-5)
    if dtype == torch.float32:
        assert torch.allclose(out, ref, rtol=2e-6, atol=1e-6)
    else:   # 1 bf16 ulp where the fp32 row sum rounds the other way
        assert torch.allclose(out.float(), ref.float(), rtol=2 ** -7, atol=1e-6)
        assert (out != ref).float().mean() < 0.01


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("rows,N,K", [(128, 3072, 1024), (128, 1024, 2816), (8, 2304, 768), (64, 1536, 4096),
                                      (200, 1024, 1024), (1, 16384, 1024), (128, 16384, 1024)])
def test_linear_bf16_tcgen05(rows, N, K):
    g = torch.Generator().manual_seed(rows + N + K)
    x = torch.randn(rows, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV, torch.bfloat16)
    out = torch.empty(rows, N, device=DEV, dtype=torch.bfloat16)
    L.check(L.lib().vlg_linear(L.ptr(x), L.ptr(w), L.ptr(out), rows, N, K, L.VLG_BF16, 0, L.stream()))
    exact = x.double() @ w.double().t()
    ref = torch.nn.functional.linear(x, w)
    # both are one rounding of the exact product to bf16 up to fp32 summation order: <= 1 bf16 ulp apart,
    # and as close to the exact result as the library GEMM is
    err_ours = (out.double() - exact).abs().max().item()
    err_ref = (ref.double() - exact).abs().max().item()
    assert err_ours <= 1.5 * err_ref + 1e-6, (err_ours, err_ref)
    assert torch.allclose(out.float(), ref.float(), rtol=2 ** -7, atol=2e-3)
    outf = torch.empty(rows, N, device=DEV, dtype=torch.float32)
    L.check(L.lib().vlg_linear(L.ptr(x), L.ptr(w), L.ptr(outf), rows, N, K, L.VLG_BF16, 1, L.stream()))
    assert torch.equal(outf, out.float())                             # `.float()` of the bf16 result, gpt.py:371


@pytest.mark.parametrize("rows,N,K", [(8, 2304, 768), (6, 512, 128), (17, 768, 2048), (128, 1024, 1024)])
def test_linear_fp32_exact_path(rows, N, K):
    g = torch.Generator().manual_seed(rows + N + K)
    x = torch.randn(rows, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.05).to(DEV)
    out = torch.empty(rows, N, device=DEV)
    L.check(L.lib().vlg_linear(L.ptr(x), L.ptr(w), L.ptr(out), rows, N, K, L.VLG_F32, 0, L.stream()))
    exact = x.double() @ w.double().t()
    assert (out.double() - exact).abs().max().item() < 2e-5           # fp32 FMA accuracy, no TF32
    out2 = torch.empty_like(out)
    L.check(L.lib().vlg_linear(L.ptr(x), L.ptr(w), L.ptr(out2), rows, N, K, L.VLG_F32, 0, L.stream()))
    assert torch.equal(out, out2)                                     # fixed summation order


# ------------------------------------------------------------------------------------------ attention
def _attn_case(rows, H, S_max, pos, dtype, masked, seed=0):
    D = 64
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(rows, H, D, generator=g).to(DEV, dtype)
    k = torch.randn(rows, H, S_max, D, generator=g).to(DEV, dtype)
    v = torch.randn(rows, H, S_max, D, generator=g).to(DEV, dtype)
    T = 0
    mask = None
    vis = torch.zeros(rows, S_max, dtype=torch.bool)
    vis[:, :pos + 1] = True
    if masked:
        T = min(120, pos + 1)
        m = (torch.rand(rows, T, generator=g) > 0.5)
        m[0] = False                                                   # a fully padded prefix
        vis[:, :T] &= m
        vis[:, pos] = True                                             # diagonal forced on
        mask = m.to(torch.uint8).to(DEV).contiguous()
    ref = O.sdpa_math(q.unsqueeze(2), k, v, vis.to(DEV)[:, None, None, :]).reshape(rows, H * D)
    return q, k, v, mask, T, ref


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("rows,H,S_max,pos,masked", [
    (128, 16, 584, 0, False), (128, 16, 584, 31, False), (128, 16, 584, 32, False), (128, 16, 584, 583, False),
    (8, 12, 264, 200, False), (2, 20, 5240, 5239, True), (16, 20, 1000, 130, True), (6, 2, 136, 119, True),
    (3, 2, 136, 57, True)])
def test_attn_decode_bf16(impl, rows, H, S_max, pos, masked):
    q, k, v, mask, T, ref = _attn_case(rows, H, S_max, pos, torch.bfloat16, masked)
    out = torch.empty(rows, H * 64, device=DEV, dtype=torch.bfloat16)
    nb = L.lib().vlg_attn_scratch_bytes(rows, H, 64, S_max)
    scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
    L.check(L.lib().vlg_attn_decode(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), rows, H, 64, S_max, pos, L.ptr(mask), T,
                                    L.VLG_BF16, impl, L.ptr(scratch), nb, L.stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert torch.allclose(out.float(), ref.float(), rtol=2 ** -7, atol=2e-3), (out.float() - ref.float()).abs().max()


@pytest.mark.parametrize("rows,H,S_max,pos,masked", [(8, 12, 264, 263, False), (6, 2, 136, 125, True),
                                                     (4, 2, 40, 0, False)])
def test_attn_decode_fp32(rows, H, S_max, pos, masked):
    q, k, v, mask, T, ref = _attn_case(rows, H, S_max, pos, torch.float32, masked)
    out = torch.empty(rows, H * 64, device=DEV)
    nb = L.lib().vlg_attn_scratch_bytes(rows, H, 64, S_max)
    scratch = torch.empty(nb, dtype=torch.uint8, device=DEV)
    L.check(L.lib().vlg_attn_decode(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(out), rows, H, 64, S_max, pos, L.ptr(mask), T,
                                    L.VLG_F32, 1, L.ptr(scratch), nb, L.stream()))
    assert torch.allclose(out, ref, rtol=1e-5, atol=2e-6)


# ------------------------------------------------------------------------------------------ sampling
CASES = [("k50", dict(top_k=50)), ("k1", dict(top_k=1)), ("p80", dict(top_p=0.8)),
         ("k100_p90_t07", dict(top_k=100, top_p=0.9, temperature=0.7)), ("plain", dict()), ("k2000", dict(top_k=2000)),
         ("t0", dict(temperature=0.0, top_k=5))]


@pytest.mark.parametrize("case,kw", CASES)
def test_sample_matches_reference_fixture(golden, case, kw):
    f = golden("sampling.npz")
    logits = torch.from_numpy(f["logits"]).to(DEV)
    idx, probs = G.sample(logits.clone(), sample_logits=False, **kw)
    ref_p = torch.from_numpy(f[case + ".probs"]).to(DEV)
    assert idx.dtype == torch.int64 and idx.shape == (5, 1)
    assert torch.equal(probs == 0, ref_p == 0), "filtered set differs"
    assert torch.allclose(probs, ref_p, rtol=2e-5, atol=1e-9)
    if case != "t0":      # at temperature 1e-5 several entries saturate to the same probability
        assert np.array_equal(idx.cpu().numpy(), f[case + ".idx"])


def test_top_k_top_p_filtering_in_place(golden):
    f = golden("sampling.npz")
    logits = torch.from_numpy(f["logits"])[:, 0].contiguous()
    ref = O.top_k_top_p_filter(logits.clone(), top_k=100, top_p=0.9)
    got = G.top_k_top_p_filtering(logits.clone().to(DEV), top_k=100, top_p=0.9)
    assert torch.equal(got.cpu(), ref)


def test_cfg_mix_topk_and_noise_driven_draws():
    """Full-size row: V=16384, top-k 2000, CFG 4.0, sampled with caller-supplied Exp(1) noise -> same tokens as
    the oracle's argmax(probs / noise) (generate.py:95-97, 57-66; multinomial's single-draw form)."""
    g = torch.Generator().manual_seed(3)
    B, V = 64, 16384
    logits = (torch.randn(2 * B, 1, V, generator=g) * 2).bfloat16().float()   # bf16-valued logits: ties are common
    noise = torch.empty(B, V).exponential_(1.0, generator=g)
    mixed = O.cfg_mix(logits, 4.0)
    ref_idx, ref_p = O.sample(mixed.clone(), temperature=1.0, top_k=2000, top_p=1.0, sample_logits=True, noise=noise)
    idx, probs = G._run_sampler(logits[:, 0].contiguous().to(DEV), 2 * B, B, 4.0, 1, 1.0, 2000, 1.0, True,
                                noise=noise.to(DEV))
    assert torch.equal(probs.cpu() == 0, ref_p == 0)
    assert int((probs > 0).sum(dim=1).min()) >= 2000                  # ties at the threshold survive
    assert torch.allclose(probs.cpu(), ref_p, rtol=2e-5, atol=1e-10)
    assert (idx.cpu() == ref_idx).float().mean() >= 0.98              # a draw can flip only on a ~1e-7 relative near-tie
    # unguided half (cfg_interval passed): conditional logits only
    idx_c, probs_c = G._run_sampler(logits[:, 0].contiguous().to(DEV), 2 * B, B, 4.0, 0, 1.0, 0, 1.0, False)
    assert torch.equal(idx_c.cpu().flatten(), logits[:B, 0].argmax(dim=-1))


def test_in_kernel_philox_draws_follow_the_distribution():
    V, B = 64, 2048
    p = torch.softmax(torch.randn(V, generator=torch.Generator().manual_seed(0)) * 1.5, dim=0)
    logits = torch.log(p).repeat(B, 1).to(DEV).contiguous()
    a, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=11)
    b, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=11)
    c, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)                # reproducible per seed
    freq = torch.bincount(a.flatten().cpu(), minlength=V).double() / B
    chi2 = float((((freq - p.double()) ** 2) / p.double()).sum() * B)
    assert chi2 < 130, chi2                                           # 63 dof: P(chi2 > 130) ~ 1e-6


# ---- the kernel branch the in-library decode loop runs (probs_out == NULL, noise == NULL, sample_logits == 1:
#      4-wide Philox counters, log-domain race without the softmax denominator) -- sample.cu, `exact == false`
def _chi2(tokens, p, n):
    freq = torch.bincount(tokens.flatten().cpu(), minlength=p.numel()).double()
    keep = p > 0
    assert int(freq[~keep].sum()) == 0, "a filtered-out token was drawn"
    e = p[keep].double() * n
    return float((((freq[keep] - e) ** 2) / e).sum()), int(keep.sum()) - 1


def test_loop_branch_draws_follow_the_distribution():
    V, B = 64, 4096
    p = torch.softmax(torch.randn(V, generator=torch.Generator().manual_seed(0)) * 1.5, dim=0)
    logits = torch.log(p).repeat(B, 1).to(DEV).contiguous()
    a, pa = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=11, want_probs=False)
    b, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=11, want_probs=False)
    c, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=12, want_probs=False)
    d, _ = G._run_sampler(logits, B, B, 1.0, 0, 1.0, 0, 1.0, True, seed=11, want_probs=False, offset=1)
    assert pa is None
    assert torch.equal(a, b) and not torch.equal(a, c) and not torch.equal(a, d)
    chi2, dof = _chi2(a, p, B)
    assert chi2 < 130, (chi2, dof)                                    # 63 dof: P(chi2 > 130) ~ 1e-6
    # rows (CTAs) and steps (offsets) are independent streams: the same token in two of them happens with prob sum(p^2)
    same = float(p.pow(2).sum())
    for x, y in ((a, d), (a[:-1], a[1:])):
        frac = float((x == y).float().mean())
        assert abs(frac - same) < 6 * (same * (1 - same) / B) ** 0.5, (frac, same)


def test_loop_branch_full_size_topk_cfg():
    """cfg-2 sampler shape: V = 16384, CFG 4.0 over 2B rows, top-k 2000: the loop branch draws from the filtered softmax
    and never picks a removed token.  Chi-square over probability-rank bins ([0], [1], [2,4), [4,8) ... [1024, 2000)) pooled
    over the 64 rows x 128 steps, so that every cell expects hundreds of draws."""
    g = torch.Generator().manual_seed(5)
    B, V, reps = 64, 16384, 128
    logits = (torch.randn(2 * B, 1, V, generator=g) * 0.5).bfloat16().float()
    _, ref_p = O.sample(O.cfg_mix(logits, 4.0).clone(), temperature=1.0, top_k=2000, top_p=1.0, sample_logits=False)
    dl = logits[:, 0].contiguous().to(DEV)
    draws = torch.stack([G._run_sampler(dl, 2 * B, B, 4.0, 1, 1.0, 2000, 1.0, True, seed=3, want_probs=False, offset=s)[0]
                         .flatten().cpu() for s in range(reps)], dim=1)          # [B, reps]
    assert bool((ref_p.gather(1, draws) > 0).all()), "a token outside the top-k set was drawn"
    order = ref_p.argsort(dim=1, descending=True, stable=True)
    rank = torch.empty_like(order)
    rank.scatter_(1, order, torch.arange(V).expand(B, V))
    edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, V]
    psort = ref_p.gather(1, order).double()
    drank = rank.gather(1, draws)
    chi2 = 0.0
    for lo, hi in zip(edges[:-1], edges[1:]):
        e = float(psort[:, lo:hi].sum()) * reps
        c = float(((drank >= lo) & (drank < hi)).sum())
        assert e > 50, (lo, hi, e)
        chi2 += (c - e) ** 2 / e
    assert chi2 < 45, chi2                                             # 11 dof: P(chi2 > 45) ~ 5e-6


def test_loop_branch_equals_api_branch_for_the_same_seed_and_step():
    """Same Philox counters in both branches: argmax(p / E) (IEEE divisions, API) and argmax((x - max) - log E)
    (decode loop, __logf) pick the same token except on near-ties of the two largest scores."""
    g = torch.Generator().manual_seed(9)
    B, V = 64, 16384
    dl = (torch.randn(2 * B, V, generator=g) * 2).bfloat16().float().to(DEV)
    agree = n = 0
    for step in range(24):
        a, _ = G._run_sampler(dl, 2 * B, B, 4.0, 1, 1.0, 2000, 1.0, True, seed=77, want_probs=True, offset=step)
        b, _ = G._run_sampler(dl, 2 * B, B, 4.0, 1, 1.0, 2000, 1.0, True, seed=77, want_probs=False, offset=step)
        agree += int((a == b).sum())
        n += B
    assert agree >= n - 3, (agree, n)      # __logf is good to ~1e-6 relative: a flip needs a ~1e-6 near-tie of the top 2


@pytest.mark.parametrize("rows,N,K,bnp", [(128, 256, 256, 128), (128, 1024, 1024, 128), (128, 384, 512, 64), (96, 256, 192, 128), (16, 128, 64, 64)])
def test_pair_gemm_cta_group_2(rows, N, K, bnp):
    """The CTA-pair building block (tcgen05.mma.cta_group::2, 2-SM TMA loads on the leader's barrier, multicast commits, the
    2x2 accumulator layout) against a plain fp32 matmul of the same bf16 operands."""
    from vlg import _lib as L
    g = torch.Generator().manual_seed(rows * 7 + N)
    x = torch.randn(rows, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.1).to(DEV, torch.bfloat16)
    out = torch.full((rows, N), float("nan"), device=DEV, dtype=torch.float32)
    L.check(L.lib().vlg_debug_pair_gemm(L.ptr(x), L.ptr(w), L.ptr(out), rows, N, K, bnp, L.stream()))
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    assert torch.isfinite(out).all()
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-3), f"max err {(out - ref).abs().max().item()}"
