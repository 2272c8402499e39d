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
#include <stdlib.h>

#include "common.cuh"

namespace vlg {

// normalise the codebook rows and precompute |e_j|^2 (F.normalize eps = 1e-12, vq_model.py:224); *e2max = max_j |e_j|^2
// en2 / h2 (optional): the layout vq_score_kernel streams with bulk copies -- every component twice ([j][C][2]) and
// -|e_j|^2 / 2 twice ([j][2]), padded to n_pad entries with codes that can never win (zeros, -inf)
template <int C>
__global__ void vq_prep_kernel(const float* __restrict__ cb, float* __restrict__ en, float* __restrict__ e2, int n_e,
                               int l2_norm, float* __restrict__ e2max, float* __restrict__ en2, float* __restrict__ h2, int n_pad) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_e) {
    if (en2 != nullptr && j < n_pad) {
#pragma unroll
      for (int c = 0; c < C; ++c) reinterpret_cast<float2*>(en2)[(size_t)j * C + c] = make_float2(0.f, 0.f);
      reinterpret_cast<float2*>(h2)[j] = make_float2(-INFINITY, -INFINITY);
    }
    return;
  }
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
  if (e2max) atomicMax(reinterpret_cast<int*>(e2max), __float_as_int(s2));   // s2 >= 0: int order == float order
  if (en2 != nullptr) {
#pragma unroll
    for (int c = 0; c < C; ++c) reinterpret_cast<float2*>(en2)[(size_t)j * C + c] = make_float2(e[c], e[c]);
    const float h = -0.5f * s2;
    reinterpret_cast<float2*>(h2)[j] = make_float2(h, h);
  }
}


// ---------------------------------------------------------------------------------------------- fast path (cfg-5 shape)
// The exact kernel below spends 13 issue slots per (latent, code) pair -- 8 for the dot product, 5 for the reference's
// d = (|z|^2 + |e|^2) - 2 z.e and the running argmin -- and the FP32 pipe sees 10 of them.  This kernel does the dot
// products alone, two latents per instruction (fma.rn.f32x2 -> FFMA2), in the score domain
//     s_j = z.e_j - |e_j|^2 / 2        (argmax_j s_j == argmin_j d_j in exact arithmetic),
// keeps, per latent, the best score / its index / the runner-up score (a 3-input max per two codes, the index only on the
// rare new record), and proves the result instead of trusting it: the reference's fp32 evaluation of d_j can disagree with
// the score order only where two scores are closer than the rounding noise of both formulas, so a latent whose runner-up
// is within `eps` of its best (about 3 in 10 000 at cfg-5) is re-done by vq_recheck_kernel with the reference's rounding
// sequence over all codes (same arithmetic as vq_argmin_kernel: bit-identical indices by construction).
//   noise budget (unit-norm rows, |z| = |e| = 1): reference d_j: ulp(2)/2 + ulp(d)/2 + 2 * 8-term dot error <= 1.3e-6,
//   scores: 8-term fma chain <= 5e-7 each; eps = 4e-6 * (|z| + |e|max)^2 covers twice their sum.
constexpr int kVq2Threads = 128;
constexpr int kVq2Chunk = 512;     // codes per shared-memory buffer; two buffers, the next chunk lands (one bulk copy of the
                                   // pre-duplicated codebook, issued by one thread) while this one is multiplied: a block no
                                   // longer stops for ~10 % of its time to fill 72 KB with 128 threads
constexpr int kVq2Pad = 2;         // entries copied past a chunk: the one-code look-ahead reads the next chunk's first code
constexpr int kVq2BufE = (kVq2Chunk + kVq2Pad) * 8 * 2;   // floats per buffer, e part
constexpr int kVq2BufH = (kVq2Chunk + kVq2Pad) * 2;       // floats per buffer, h part
constexpr int vq2_padded(int n_e) { return (n_e + kVq2Chunk - 1) / kVq2Chunk * kVq2Chunk + kVq2Pad; }

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 1D bulk copy global -> shared (bytes a multiple of 16), completes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// LP latent pairs per thread (2 * LP latents), G codes whose scores stay in registers between index look-ups; C = 8 only.
template <int LP, int G, bool SPLIT = false>
__global__ void __launch_bounds__(kVq2Threads, LP <= 4 ? 3 : 1)
vq_score_kernel(const float* __restrict__ z, const float* __restrict__ en, const float* __restrict__ e2,
                const float* __restrict__ e2max, const float* __restrict__ en2, const float* __restrict__ h2,
                int64_t* __restrict__ idx_out, float* __restrict__ zq,
                unsigned* __restrict__ amb_count, int* __restrict__ amb_list, int64_t N, int64_t HW, int n_e, int l2_norm) {
  constexpr int C = 8, L = 2 * LP;
  extern __shared__ __align__(128) float sm[];
  // two buffers of [chunk + pad][C][2] (every component twice, so one 64-bit load feeds both halves of an FFMA2) and
  // [chunk + pad][2] (-|e_j|^2 / 2, twice), then the two "landed" mbarriers
  float* s_e_buf = sm;
  float* s_h_buf = sm + 2 * kVq2BufE;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(sm + 2 * kVq2BufE + 2 * kVq2BufH);
  const int n_chunks = (n_e + kVq2Chunk - 1) / kVq2Chunk;
  auto fetch = [&](int chunk) {                 // one thread: both parts of a chunk, one mbarrier
    const int b = chunk & 1;
    mbar_expect_tx(&s_full[b], (kVq2BufE + kVq2BufH) * 4);
    bulk_g2s(s_e_buf + b * kVq2BufE, en2 + (size_t)chunk * kVq2Chunk * C * 2, kVq2BufE * 4, &s_full[b]);
    bulk_g2s(s_h_buf + b * kVq2BufH, h2 + (size_t)chunk * kVq2Chunk * 2, kVq2BufH * 4, &s_full[b]);
  };
  if (threadIdx.x == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    fence_mbar_init();
    fetch(0);
  }
  __syncthreads();

  const int64_t base = (int64_t)blockIdx.x * (kVq2Threads * L) + threadIdx.x;
  unsigned long long zp[LP][C];             // latent pair p = latents (base + (2p) * T, base + (2p+1) * T)
  float z2[L];
#pragma unroll
  for (int p = 0; p < LP; ++p) {
    float zr[2][C];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t n = base + (int64_t)(2 * p + h) * kVq2Threads;
      const bool ok = n < N;
      const int64_t b = ok ? n / HW : 0, hw = ok ? n % HW : 0;
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        zr[h][c] = ok ? z[(b * C + c) * HW + hw] : 0.f;
        ss = __fadd_rn(ss, __fmul_rn(zr[h][c], zr[h][c]));
      }
      if (l2_norm) {
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        ss = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          zr[h][c] = __fdiv_rn(zr[h][c], nrm);
          ss = __fadd_rn(ss, __fmul_rn(zr[h][c], zr[h][c]));
        }
      }
      z2[2 * p + h] = ss;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) zp[p][c] = pack2(zr[0][c], zr[1][c]);
  }

  float best[L], second[L];
  int bgrp[L];                  // first code of the group of G codes that holds the best score so far
#pragma unroll
  for (int l = 0; l < L; ++l) { best[l] = -INFINITY; second[l] = -INFINITY; bgrp[l] = 0; }

  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int c0 = chunk * kVq2Chunk;
    // the other buffer was last read a chunk ago (everyone passed the barrier that closed it): refill it while this one is used
    if (threadIdx.x == 0 && chunk + 1 < n_chunks) fetch(chunk + 1);
    while (!mbar_try_wait(&s_full[chunk & 1], (chunk >> 1) & 1)) {}
    const float* s_e = s_e_buf + (chunk & 1) * kVq2BufE;
    const float* s_h = s_h_buf + (chunk & 1) * kVq2BufH;
    // e-values of code j + 1 are requested while code j is being multiplied (one code = 4 * LP FFMA2 issue slots, longer
    // than a shared-memory round trip); a buffer holds kVq2Pad codes past its chunk so the look-ahead never leaves it
    const ulonglong2* ep0 = reinterpret_cast<const ulonglong2*>(s_e);
    ulonglong2 n01 = ep0[0], n23 = ep0[1], n45 = ep0[2], n67 = ep0[3];
    unsigned long long nh = reinterpret_cast<const unsigned long long*>(s_h)[0];
#pragma unroll 1
    for (int j0 = 0; j0 < kVq2Chunk; j0 += G) {
      unsigned long long sc[LP][G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const ulonglong2 e01 = n01, e23 = n23, e45 = n45, e67 = n67;
        const unsigned long long h = nh;
        const ulonglong2* ep = reinterpret_cast<const ulonglong2*>(s_e + (size_t)(j0 + j + 1) * C * 2);   // warp-wide broadcasts
        n01 = ep[0]; n23 = ep[1]; n45 = ep[2]; n67 = ep[3];
        nh = reinterpret_cast<const unsigned long long*>(s_h)[j0 + j + 1];
#pragma unroll
        for (int p = 0; p < LP; ++p) {
          if constexpr (SPLIT) {     // two 4-deep chains and an add: twice the independent work per code for one more pipe slot in eight
            unsigned long long a = ffma2(zp[p][0], e01.x, h), b = fmul2(zp[p][4], e45.x);
            a = ffma2(zp[p][1], e01.y, a); b = ffma2(zp[p][5], e45.y, b);
            a = ffma2(zp[p][2], e23.x, a); b = ffma2(zp[p][6], e67.x, b);
            a = ffma2(zp[p][3], e23.y, a); b = ffma2(zp[p][7], e67.y, b);
            sc[p][j] = fadd2(a, b);
          } else {
            unsigned long long a = ffma2(zp[p][0], e01.x, h);
            a = ffma2(zp[p][1], e01.y, a);
            a = ffma2(zp[p][2], e23.x, a);
            a = ffma2(zp[p][3], e23.y, a);
            a = ffma2(zp[p][4], e45.x, a);
            a = ffma2(zp[p][5], e45.y, a);
            a = ffma2(zp[p][6], e67.x, a);
            sc[p][j] = ffma2(zp[p][7], e67.y, a);
          }
        }
      }
#pragma unroll
      for (int l = 0; l < L; ++l) {
        float s[G];
#pragma unroll
        for (int j = 0; j < G; ++j) s[j] = (l & 1) ? hi32(sc[l >> 1][j]) : lo32(sc[l >> 1][j]);
        float m;
        if constexpr (G == 8) m = max3f(max3f(s[0], s[1], s[2]), max3f(s[3], s[4], s[5]), fmaxf(s[6], s[7]));
        else m = fmaxf(max3f(s[0], s[1], s[2]), s[3]);
        // Branch-free record keeping at GROUP granularity (which code of the group it was is settled once, after the scan):
        // a new record demotes the old best to runner-up, anything else competes for runner-up; equal maxima keep the earlier
        // group (first index) and leave a zero gap, which sends the latent to the exact recheck.  Eight independent 5-instruction
        // chains per iteration instead of eight convergence-barrier regions (ncu: those held the FMA pipe at 55 %).
        second[l] = fmaxf(second[l], fminf(m, best[l]));
        bgrp[l] = m > best[l] ? c0 + j0 : bgrp[l];
        best[l] = fmaxf(best[l], m);
      }
    }
    __syncthreads();                            // this buffer may be refilled from now on
  }

  // which code of the winning group, and the runner-up inside it: the group's scores again, with the loop's arithmetic
  // (an FFMA2 lane is an IEEE fma, the scalar chain below reproduces it bit for bit)
  int bidx[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    float zl[C];
#pragma unroll
    for (int c = 0; c < C; ++c) zl[c] = (l & 1) ? hi32(zp[l >> 1][c]) : lo32(zp[l >> 1][c]);
    int arg = -1;
    float run = second[l];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int code = bgrp[l] + j;
      float a = -INFINITY;
      if (code < n_e) {
        const float* e = en + (size_t)code * C;
        a = -0.5f * e2[code];
        if constexpr (SPLIT) {
          float b = __fmul_rn(zl[4], e[4]);
#pragma unroll
          for (int c = 0; c < 4; ++c) a = __fmaf_rn(zl[c], e[c], a);
#pragma unroll
          for (int c = 5; c < 8; ++c) b = __fmaf_rn(zl[c], e[c], b);
          a = __fadd_rn(a, b);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) a = __fmaf_rn(zl[c], e[c], a);
        }
      }
      if (arg < 0 && a == best[l]) arg = j;                  // first index among equals
      else run = fmaxf(run, a);
    }
    second[l] = arg < 0 ? best[l] : run;                     // no match (NaN scores): zero gap, the recheck decides
    bidx[l] = bgrp[l] + (arg < 0 ? 0 : arg);
  }

  const float emax = e2max ? *e2max : 1.f;
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const int64_t n = base + (int64_t)l * kVq2Threads;
    if (n >= N) continue;
    const float mag = sqrtf(z2[l]) + sqrtf(emax);
    const float eps = 4e-6f * fmaxf(mag * mag, 1e-30f);
    if (!(best[l] - second[l] > eps)) {      // runner-up inside the rounding noise (or NaN scores): settle it exactly
      const unsigned slot = atomicAdd(amb_count, 1u);
      amb_list[slot] = (int)n;               // amb_list holds N entries: cannot overflow
      continue;
    }
    idx_out[n] = bidx[l];
    if (zq) {
      const int64_t b = n / HW, hw = n % HW;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float e = en[(size_t)bidx[l] * C + c];
        const float zv = (l & 1) ? hi32(zp[l >> 1][c]) : lo32(zp[l >> 1][c]);
        zq[(b * C + c) * HW + hw] = __fadd_rn(zv, __fsub_rn(e, zv));
      }
    }
  }
}

// One CTA per ambiguous latent: the reference's rounding sequence for d over all codes (same arithmetic as
// vq_argmin_kernel), the 256 threads stride the codebook (each scans its codes in increasing order, strict '<': first index
// on ties), minimum with the lowest index on ties across the block.  (One WARP per latent, the first version, left ~300 of
// 2368 warps busy for 512 dependent iterations each: 281 us, 4 % of the whole index search.)
template <int C>
__global__ void __launch_bounds__(256)
vq_recheck_kernel(const float* __restrict__ z, const float* __restrict__ en, const float* __restrict__ e2,
                  int64_t* __restrict__ idx_out, float* __restrict__ zq, const unsigned* __restrict__ amb_count,
                  const int* __restrict__ amb_list, int64_t HW, int n_e, int l2_norm) {
  __shared__ float s_best[8];
  __shared__ int s_bi[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned total = *amb_count;
  for (unsigned w = blockIdx.x; w < total; w += gridDim.x) {
    const int64_t n = amb_list[w];
    const int64_t b = n / HW, hw = n % HW;
    float zr[C];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      zr[c] = z[(b * C + c) * HW + hw];
      ss = __fadd_rn(ss, __fmul_rn(zr[c], zr[c]));
    }
    if (l2_norm) {
      const float nrm = fmaxf(sqrtf(ss), 1e-12f);
      ss = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        zr[c] = __fdiv_rn(zr[c], nrm);
        ss = __fadd_rn(ss, __fmul_rn(zr[c], zr[c]));
      }
    }
    float best = INFINITY;
    int bi = 0x7fffffff;
#pragma unroll 4
    for (int j = threadIdx.x; j < n_e; j += 256) {
      float e[C];
      if constexpr (C % 4 == 0) {
#pragma unroll
        for (int c = 0; c < C; c += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(en + (size_t)j * C + c));
          e[c] = v.x; e[c + 1] = v.y; e[c + 2] = v.z; e[c + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) e[c] = en[(size_t)j * C + c];
      }
      float dot = __fmul_rn(zr[0], e[0]);
#pragma unroll
      for (int c = 1; c < C; ++c) dot = __fmaf_rn(zr[c], e[c], dot);
      const float d = __fmaf_rn(-2.f, dot, __fadd_rn(ss, e2[j]));
      if (d < best) { best = d; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    __syncthreads();                         // the previous item's readers of s_best / s_bi are done
    if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; }
    __syncthreads();
    best = s_best[0];
    bi = s_bi[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float ob = s_best[k];
      const int oi = s_bi[k];
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (threadIdx.x >= 32) continue;         // warp 0 writes the result
    if (bi == 0x7fffffff) bi = 0;            // all-NaN row: torch.argmin would return the first NaN; index 0 here
    if (lane == 0) idx_out[n] = bi;
    if (zq && lane < C) zq[(b * C + lane) * HW + hw] = __fadd_rn(zr[lane], __fsub_rn(en[(size_t)bi * C + lane], zr[lane]));
  }
}

// ---------------------------------------------------------------------------------------------- exact kernel (any C)
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

// get_codebook_entry + post_quant_conv (1x1): one thread per latent position holds its normalised code in registers and
// walks the O output channels; the [O, C] weight and the bias sit in shared memory and are read as warp broadcasts, the
// stores are coalesced along hw.  HBM-bound on the [B, O, HW] write.
template <int C>
__global__ void __launch_bounds__(128)
codebook_project_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb, const float* __restrict__ w,
                        const float* __restrict__ bias, float* __restrict__ out, int64_t N, int64_t HW, int n_e, int O,
                        int l2_norm) {
  extern __shared__ __align__(16) float sm[];
  float* s_w = sm;            // [O][C]
  float* s_b = sm + O * C;    // [O]
  for (int i = threadIdx.x; i < O * C; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < O; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
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
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int c = 0; c < C; ++c) e[c] = __fdiv_rn(e[c], nrm);
  }
  const int64_t b = n / HW, hw = n - b * HW;
  float* dst = out + b * O * HW + hw;
  for (int o = 0; o < O; ++o) {
    float acc = s_b[o];
#pragma unroll
    for (int c = 0; c < C; ++c) acc = __fmaf_rn(s_w[o * C + c], e[c], acc);
    dst[(int64_t)o * HW] = acc;
  }
}

template <int C, int L>
static int launch_vq(const float* z, const float* cb, int64_t* idx, float* zq, float* scratch, int64_t N, int64_t HW,
                     int n_e, int l2_norm, cudaStream_t st) {
  float* en = scratch;
  float* e2 = scratch + (size_t)n_e * C;
  vq_prep_kernel<C><<<(n_e + 255) / 256, 256, 0, st>>>(cb, en, e2, n_e, l2_norm, nullptr, nullptr, nullptr, 0);
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

namespace vlg {
// C = 8 with the work area of vlg_vq_scratch_floats(): score-domain FFMA2 kernel + exact recheck of the ambiguous latents
static int launch_vq_fast(const float* z, const float* cb, int64_t* idx, float* zq, float* scratch, int64_t N, int64_t HW,
                          int n_e, int l2_norm, cudaStream_t st) {
  constexpr int C = 8;
  float* en = scratch;
  float* e2 = scratch + (size_t)n_e * C;
  float* e2max = e2 + n_e;                                         // 1 float
  unsigned* amb_count = reinterpret_cast<unsigned*>(e2max + 1);    // 1 word
  const int n_pad = vq2_padded(n_e);
  float* en2 = e2max + 4;                                          // [n_pad][C][2], 16-byte aligned (the scratch base is)
  float* h2 = en2 + (size_t)n_pad * C * 2;                         // [n_pad][2]
  int* amb_list = reinterpret_cast<int*>(h2 + (size_t)n_pad * 2);  // N words
  VLG_REQUIRE((reinterpret_cast<uintptr_t>(en2) & 15) == 0, "vlg_vq_argmin: scratch must be 16-byte aligned");
  VLG_CUDA(cudaMemsetAsync(e2max, 0, 2 * sizeof(float), st));
  vq_prep_kernel<C><<<(n_pad + 255) / 256, 256, 0, st>>>(cb, en, e2, n_e, l2_norm, e2max, en2, h2, n_pad);
  VLG_LAUNCH_CHECK();
  const size_t smem = (size_t)(2 * kVq2BufE + 2 * kVq2BufH) * sizeof(float) + 2 * sizeof(uint64_t);
  // (latent pairs per thread, codes per register group): measured on B200 by tools/vq_bench.py; VLG_VQ_VARIANT for A/B
  static const int variant = [] { const char* e = getenv("VLG_VQ_VARIANT"); return e ? atoi(e) : 0; }();
  auto launch = [&](auto kern, int lp) -> int {
    VLG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per_block = (int64_t)kVq2Threads * 2 * lp;
    kern<<<(unsigned)((N + per_block - 1) / per_block), kVq2Threads, smem, st>>>(z, en, e2, e2max, en2, h2, idx, zq, amb_count, amb_list,
                                                                                 N, HW, n_e, l2_norm);
    VLG_LAUNCH_CHECK();
    return VLG_OK;
  };
  int rc;
  switch (variant) {
    case 1: rc = launch(vq_score_kernel<6, 4>, 6); break;
    case 2: rc = launch(vq_score_kernel<3, 4>, 3); break;
    case 3: rc = launch(vq_score_kernel<3, 8>, 3); break;
    case 4: rc = launch(vq_score_kernel<2, 8>, 2); break;
    case 5: rc = launch(vq_score_kernel<4, 8>, 4); break;
    case 6: rc = launch(vq_score_kernel<4, 4, true>, 4); break;
    case 7: rc = launch(vq_score_kernel<4, 8, true>, 4); break;
    case 8: rc = launch(vq_score_kernel<3, 8, true>, 3); break;
    default: rc = launch(vq_score_kernel<4, 4>, 4); break;   // 6.52 ms / 1M latents ((4, 8): 6.55, spills at the 168-register cap; split chains: 6.8)
  }
  if (rc) return rc;
  vq_recheck_kernel<C><<<num_sms() * 4, 256, 0, st>>>(z, en, e2, idx, zq, amb_count, amb_list, HW, n_e, l2_norm);
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}
}  // namespace vlg

extern "C" size_t vlg_vq_scratch_floats(int64_t N, int C, int n_e) {
  // normalised codebook, |e|^2, {max |e|^2, ambiguous count, 2 spare}, the duplicated padded layout of the score kernel, list
  const size_t n_pad = (size_t)vlg::vq2_padded(n_e);
  return (size_t)n_e * (C + 1) + 4 + n_pad * ((size_t)C * 2 + 2) + (size_t)(N > 0 ? N : 0);
}

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
  static const bool exact_only = [] { const char* e = getenv("VLG_VQ_EXACT"); return e && e[0] == '1'; }();   // A/B switch
  if (C == 8 && !exact_only && N < (1ll << 31)) return launch_vq_fast(z_dev, codebook_dev, idx_dev, zq_dev, scratch_dev, N, HW, n_e, l2_norm, st);
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

extern "C" int vlg_codebook_project(const int64_t* idx_dev, const float* codebook_dev, const float* weight_dev,
                                    const float* bias_dev, float* out_dev, int64_t B, int C, int64_t HW, int n_e, int O,
                                    int l2_norm, void* stream) {
  using namespace vlg;
  const int64_t N = B * HW;
  if (N == 0) return VLG_OK;
  VLG_REQUIRE(idx_dev && codebook_dev && weight_dev && out_dev, "vlg_codebook_project: null pointer");
  VLG_REQUIRE(O > 0 && (size_t)O * (C + 1) * sizeof(float) <= 48 * 1024, "vlg_codebook_project: %d output channels do not fit shared memory", O);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((N + 127) / 128);
  const size_t smem = (size_t)O * (C + 1) * sizeof(float);
  switch (C) {
    case 4: codebook_project_kernel<4><<<blocks, 128, smem, st>>>(idx_dev, codebook_dev, weight_dev, bias_dev, out_dev, N, HW, n_e, O, l2_norm); break;
    case 8: codebook_project_kernel<8><<<blocks, 128, smem, st>>>(idx_dev, codebook_dev, weight_dev, bias_dev, out_dev, N, HW, n_e, O, l2_norm); break;
    case 16: codebook_project_kernel<16><<<blocks, 128, smem, st>>>(idx_dev, codebook_dev, weight_dev, bias_dev, out_dev, N, HW, n_e, O, l2_norm); break;
    case 32: codebook_project_kernel<32><<<blocks, 128, smem, st>>>(idx_dev, codebook_dev, weight_dev, bias_dev, out_dev, N, HW, n_e, O, l2_norm); break;
    default:
      set_error("vlg_codebook_project: codebook_embed_dim %d not supported (4, 8, 16, 32)", C);
      return VLG_ERR_ARG;
  }
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

// ---------------------------------------------------------------------------------------------- FP32 roofline probe
// Register-resident FMA loop (the VQ index search is bounded by the FP32 pipe, so its roofline denominator is
// measured on the box, not quoted): 8 independent accumulator chains per thread, 2 CTAs of 512 threads per SM.
// variant 0: scalar FFMA, variant 1: packed FFMA2 (fma.rn.f32x2).  flops = grid * 512 * iters * 16 * (variant ? 2 : 1).
namespace vlg {
template <int X2>
__global__ void __launch_bounds__(512) fp32_peak_kernel(float* out, int iters, float a, float b) {
  if (X2) {
    unsigned long long acc[8], av = pack2(a, a), bv = pack2(b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = pack2((float)threadIdx.x + i, (float)i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = ffma2(acc[i], av, bv);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += lo32(acc[i]) + hi32(acc[i]);
    if (s == 12345.678f) out[0] = s;
  } else {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (float)threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678f) out[0] = s;
  }
}
}  // namespace vlg

extern "C" int vlg_debug_fp32_fma_peak(int variant, int iters, float* out_dev, long long* flops_out, void* stream) {
  using namespace vlg;
  VLG_REQUIRE(out_dev && iters > 0, "vlg_debug_fp32_fma_peak: bad arguments");
  const int grid = num_sms() * 2;
  if (variant) fp32_peak_kernel<1><<<grid, 512, 0, (cudaStream_t)stream>>>(out_dev, iters, 0.999f, 0.001f);
  else fp32_peak_kernel<0><<<grid, 512, 0, (cudaStream_t)stream>>>(out_dev, iters, 0.999f, 0.001f);
  VLG_LAUNCH_CHECK();
  if (flops_out) *flops_out = (long long)grid * 512 * iters * 16 * (variant ? 2 : 1);
  return VLG_OK;
}
