// VQ tokenizer index search: argmin_j ( |z|^2 + |e_j|^2 - 2 z.e_j ) over the (l2-normalised) codebook,
// without ever materialising the [N, n_e] distance matrix the reference builds (vq_model.py:228-232).
//
// Bound: FP32 FMA pipe, not HBM: 2*C flops per (latent, code) pair against 40 bytes of
// traffic per thousand pairs.  Layout: each thread keeps L latents (C floats each) in registers; the
// normalised codebook is staged through shared memory in chunks and read as warp-wide broadcasts, so
// one pair costs C FMAs + 1 add + 1 fma + compare/select.
//
// Exactness: d is formed with the reference's rounding sequence -- (|z|^2 + |e|^2) rounded, then
// minus 2*dot rounded (fma(-2, dot, s) is that exactly, doubling being exact) -- and ties keep the
// lowest index, like torch.argmin.
#include "common.cuh"

namespace vlg {

// normalise the codebook rows and precompute |e_j|^2 (F.normalize eps = 1e-12, vq_model.py:224)
template <int C>
__global__ void vq_prep_kernel(const float* __restrict__ cb, float* __restrict__ en, float* __restrict__ e2, int n_e,
                               int l2_norm) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_e) return;
  float e[C];
#pragma unroll
  for (int c = 0; c < C; ++c) e[c] = cb[(size_t)j * C + c];
  if (l2_norm) {
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) ss = __fadd_rn(ss, __fmul_rn(e[c], e[c]));
    float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = __fdiv_rn(e[c], nrm);
  }
  float s2 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    s2 = __fadd_rn(s2, __fmul_rn(e[c], e[c]));
    en[(size_t)j * C + c] = e[c];
  }
  e2[j] = s2;
}

constexpr int kVqThreads = 256;
constexpr int kVqChunk = 2048;  // codes staged per shared-memory fill

template <int C, int L>
__global__ void __launch_bounds__(kVqThreads)
vq_argmin_kernel(const float* __restrict__ z, const float* __restrict__ en, const float* __restrict__ e2,
                 int64_t* __restrict__ idx_out, float* __restrict__ zq, int64_t N, int64_t HW, int n_e, int l2_norm) {
  extern __shared__ __align__(16) float sm[];
  float* s_e = sm;                   // [kVqChunk][C]
  float* s_e2 = sm + kVqChunk * C;   // [kVqChunk]

  // latents of this thread: n = base + l * kVqThreads (coalesced across the block for every channel)
  const int64_t base = (int64_t)blockIdx.x * (kVqThreads * L) + threadIdx.x;
  float zr[L][C], z2[L], best[L];
  int bidx[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    int64_t n = base + (int64_t)l * kVqThreads;
    bool ok = n < N;
    int64_t b = ok ? n / HW : 0, hw = ok ? n % HW : 0;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      zr[l][c] = ok ? z[(b * C + c) * HW + hw] : 0.f;
      ss = __fadd_rn(ss, __fmul_rn(zr[l][c], zr[l][c]));
    }
    if (l2_norm) {
      float nrm = fmaxf(sqrtf(ss), 1e-12f);
      ss = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        zr[l][c] = __fdiv_rn(zr[l][c], nrm);
        ss = __fadd_rn(ss, __fmul_rn(zr[l][c], zr[l][c]));
      }
    }
    z2[l] = ss;
    best[l] = INFINITY;
    bidx[l] = 0;
  }

  for (int c0 = 0; c0 < n_e; c0 += kVqChunk) {
    int cn = min(kVqChunk, n_e - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < cn * C / 4; i += kVqThreads)
      reinterpret_cast<float4*>(s_e)[i] = reinterpret_cast<const float4*>(en + (size_t)c0 * C)[i];
    for (int i = threadIdx.x; i < cn; i += kVqThreads) s_e2[i] = e2[c0 + i];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cn; ++j) {
      float e[C];
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        float4 v = reinterpret_cast<const float4*>(s_e)[j * (C / 4) + q];  // warp-wide broadcast
        e[4 * q] = v.x; e[4 * q + 1] = v.y; e[4 * q + 2] = v.z; e[4 * q + 3] = v.w;
      }
      float ee = s_e2[j];
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float dot = __fmul_rn(zr[l][0], e[0]);
#pragma unroll
        for (int c = 1; c < C; ++c) dot = __fmaf_rn(zr[l][c], e[c], dot);
        float d = __fmaf_rn(-2.f, dot, __fadd_rn(z2[l], ee));
        if (d < best[l]) { best[l] = d; bidx[l] = c0 + j; }
      }
    }
  }

#pragma unroll
  for (int l = 0; l < L; ++l) {
    int64_t n = base + (int64_t)l * kVqThreads;
    if (n >= N) continue;
    idx_out[n] = bidx[l];
    if (zq) {
      int64_t b = n / HW, hw = n % HW;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float e = en[(size_t)bidx[l] * C + c];
        // straight-through form z + (z_q - z) (vq_model.py:254), evaluated like the reference
        zq[(b * C + c) * HW + hw] = __fadd_rn(zr[l][c], __fsub_rn(e, zr[l][c]));
      }
    }
  }
}

template <int C>
__global__ void codebook_gather_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb,
                                       float* __restrict__ out, int64_t N, int64_t HW, int n_e, int l2_norm,
                                       int channel_first) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t j = idx[n];
  if (j < 0 || j >= n_e) j = 0;
  float e[C];
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    e[c] = cb[j * C + c];
    ss = __fadd_rn(ss, __fmul_rn(e[c], e[c]));
  }
  if (l2_norm) {
    float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = __fdiv_rn(e[c], nrm);
  }
  int64_t b = n / HW, hw = n % HW;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    if (channel_first) out[(b * C + c) * HW + hw] = e[c];
    else out[n * C + c] = e[c];
  }
}

template <int C, int L>
static int launch_vq(const float* z, const float* cb, int64_t* idx, float* zq, float* scratch, int64_t N, int64_t HW,
                     int n_e, int l2_norm, cudaStream_t st) {
  float* en = scratch;
  float* e2 = scratch + (size_t)n_e * C;
  vq_prep_kernel<C><<<(n_e + 255) / 256, 256, 0, st>>>(cb, en, e2, n_e, l2_norm);
  VLG_LAUNCH_CHECK();
  size_t smem = (size_t)kVqChunk * (C + 1) * sizeof(float);
  auto kern = vq_argmin_kernel<C, L>;
  VLG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t per_block = (int64_t)kVqThreads * L;
  int64_t blocks = (N + per_block - 1) / per_block;
  kern<<<(unsigned)blocks, kVqThreads, smem, st>>>(z, en, e2, idx, zq, N, HW, n_e, l2_norm);
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg

extern "C" int vlg_vq_argmin(const float* z_dev, const float* codebook_dev, int64_t* idx_dev, float* zq_dev,
                             float* scratch_dev, int64_t B, int C, int64_t HW, int n_e, int l2_norm, void* stream) {
  using namespace vlg;
  VLG_REQUIRE(B >= 0 && HW >= 0 && n_e > 0, "vlg_vq_argmin: bad sizes B=%lld HW=%lld n_e=%d", (long long)B,
              (long long)HW, n_e);
  int64_t N = B * HW;
  if (N == 0) return VLG_OK;   // empty batch: nothing to do (torch hands out null pointers for empty tensors)
  VLG_REQUIRE(z_dev && codebook_dev && idx_dev && scratch_dev, "vlg_vq_argmin: null pointer");
  VLG_REQUIRE(N / ((int64_t)kVqThreads) < (1ll << 31), "vlg_vq_argmin: too many latents");
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 4: return launch_vq<4, 4>(z_dev, codebook_dev, idx_dev, zq_dev, scratch_dev, N, HW, n_e, l2_norm, st);
    case 8: return launch_vq<8, 4>(z_dev, codebook_dev, idx_dev, zq_dev, scratch_dev, N, HW, n_e, l2_norm, st);
    case 16: return launch_vq<16, 2>(z_dev, codebook_dev, idx_dev, zq_dev, scratch_dev, N, HW, n_e, l2_norm, st);
    case 32: return launch_vq<32, 1>(z_dev, codebook_dev, idx_dev, zq_dev, scratch_dev, N, HW, n_e, l2_norm, st);
    default:
      set_error("vlg_vq_argmin: codebook_embed_dim %d not supported (4, 8, 16, 32)", C);
      return VLG_ERR_ARG;
  }
}

extern "C" int vlg_codebook_gather(const int64_t* idx_dev, const float* codebook_dev, float* out_dev, int64_t B, int C,
                                   int64_t HW, int n_e, int l2_norm, int channel_first, void* stream) {
  using namespace vlg;
  int64_t N = B * HW;
  if (N == 0) return VLG_OK;
  VLG_REQUIRE(idx_dev && codebook_dev && out_dev, "vlg_codebook_gather: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned blocks = (unsigned)((N + 255) / 256);
  switch (C) {
    case 4: codebook_gather_kernel<4><<<blocks, 256, 0, st>>>(idx_dev, codebook_dev, out_dev, N, HW, n_e, l2_norm, channel_first); break;
    case 8: codebook_gather_kernel<8><<<blocks, 256, 0, st>>>(idx_dev, codebook_dev, out_dev, N, HW, n_e, l2_norm, channel_first); break;
    case 16: codebook_gather_kernel<16><<<blocks, 256, 0, st>>>(idx_dev, codebook_dev, out_dev, N, HW, n_e, l2_norm, channel_first); break;
    case 32: codebook_gather_kernel<32><<<blocks, 256, 0, st>>>(idx_dev, codebook_dev, out_dev, N, HW, n_e, l2_norm, channel_first); break;
    default:
      set_error("vlg_codebook_gather: codebook_embed_dim %d not supported (4, 8, 16, 32)", C);
      return VLG_ERR_ARG;
  }
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}
