// Diffusion-head token sampler: DiffLoss.sample of the reference (autoregressive/models/diffloss.py:35-52) -- a
// `steps`-long (100) DDPM chain per generated token, each step one pass through SimpleMLPAdaLN (diffloss.py:151-235:
// input_proj, `depth` adaLN-modulated residual MLP blocks, a modulated final layer) followed by the eps / learned-range
// posterior update of gaussian_diffusion.py:254-332, 376-420.  Eager PyTorch spends ~40 launches per diffusion step, i.e.
// ~4000 launches per token, on [rows x 1024] operands.  fp32 throughout (the shipped sampler runs --precision none,
// scripts/sample/sample_t2v_diff.bash:30).
//
// What the chain looks like once the parts that do not depend on the diffusing sample are taken out of it:
//   * the conditioning c = cond_embed(z) is fixed for the token, the timestep embedding is fixed for the model, so every
//     adaLN modulation  Linear(SiLU(t_emb[s] + c[r]))  of every step is known before the chain starts: ONE fp32 GEMM
//     [steps*rows, W] x [W, (3*depth+2) W] per token (dh_sgemm_kernel, CUDA-core fp32) replaces depth+1 of the dependent
//     layers of every step and takes 2/3 of the weights out of the loop;
//   * what is left per step is a dependent chain of 2*depth [W x W] layers and the [2C x W] output layer on `rows` vectors:
//     no arithmetic to speak of (6 MFLOP per row), only latency.  dh_sample_kernel keeps those weights RESIDENT IN SHARED
//     MEMORY for the whole chain -- 2*depth*W*W*4 B = 25 MB at W = 1024, depth = 3, spread over 128 CTAs x 192 KB, 8
//     output columns each -- so that a layer is: read the [rows x W] activation from L2, LayerNorm + modulate in
//     registers (one row per warp at a time, lane-strided), warp-wide dot products against shared memory,
//     store 8 columns, grid barrier.  2*depth + 1 barriers per diffusion step, nothing else leaves the SM; the posterior
//     update is done redundantly by every CTA from the same [rows x 2C] output (deterministic: same inputs, same code).
// One cooperative launch per token (the driver guarantees the grid is co-resident, which the barriers need).
#include <vector>

#include "kernels.h"
#include "philox.cuh"

namespace vlg {

constexpr int kDhThreads = 256;
constexpr int kDhWarps = kDhThreads / 32;
constexpr int kDhMaxC = 16;           // latent channels (the CausalVideoVAE of the reference: 8)
constexpr int kDhMaxRows = 16;
constexpr int kDhMaxDepth = 8;
constexpr int kDhTimeFreq = 256;      // TimestepEmbedder.frequency_embedding_size (diffloss.py:64)

struct DhBlock { const float *ln_w, *ln_b, *w1, *b1, *w2, *b2; };

struct DhParams {
  int R, C, W, depth, steps;
  int cpc, G;                // column CTAs: G = W / cpc CTAs own cpc output columns of every W x W layer; CTAs [G, G + 2C) own
                             // one column of the output layer each
  int n_mod;                 // (3 * depth + 2) * W
  float temperature, cfg_scale;
  int use_cfg;               // forward_with_cfg (diffloss.py:237-247): rows [0, R/2) conditional, [R/2, R) unconditional
  const float *in_w, *in_b;  // input_proj [W, C], [W]
  DhBlock blk[kDhMaxDepth];
  const float *final_w, *final_b;   // final_layer.linear [2C, W], [2C]
  const float* ada_b;        // [n_mod]: adaLN biases, block b at [3 W b, +3 W) as (shift | scale | gate), final at [3 W depth, +2 W)
  const float* mod;          // [steps * R, n_mod] = SiLU(t_emb[s] + c[r]) . W_ada^T (bias not yet added), row s * R + r
  const float* tab;          // [6][steps]: sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1,
                             //             posterior_mean_coef2, posterior_log_variance_clipped, log(betas)
  const float* noise;        // [R, C] start sample, or null -> Philox
  const float* step_noise;   // [steps, R, C] draw of loop index i, or null -> Philox
  unsigned long long seed, offset;
  float *hbuf, *vbuf, *obuf; // [R, W] residual stream, [R, W] hidden of the block MLP, [R, 2C] model output
  float* out;                // [R, C]
  unsigned* bar;             // [0] barrier arrivals, [1] exit arrivals; zero between launches
  volatile unsigned* err;    // mapped host word: a barrier gave up waiting
};

__device__ __forceinline__ float4 dh_ldcg4(const float* p) {
  float4 r;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float dh_ldcg(const float* p) {
  float r;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float dh_silu(float x) { return __fdiv_rn(x, 1.0f + expf(-x)); }

// Same protocol as the layer-tail kernel's barrier (layer_tail.cu): release add, acquire poll on one L2 word; the wait is
// bounded and ends in a host-visible error word, never in a trap.
__device__ __forceinline__ void dh_grid_barrier(unsigned* ctr, unsigned& target, volatile unsigned* err) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned seen, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
      if (seen >= target) break;
      if ((++spins & 1023u) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        if (err != nullptr && *err != 0) break;
        if (now - t0 > 2000000000ull) {
          if (err != nullptr) { *err = 1u; __threadfence_system(); }
          break;
        }
      }
    }
  }
  __syncthreads();
}

// standard normal from Philox4x32-10 (Box-Muller on two 24-bit uniforms); counter = (element, loop index, offset)
__device__ __forceinline__ float dh_normal(unsigned long long seed, unsigned long long off, uint32_t step, uint32_t elem) {
  const uint4 r = philox4x32_10(make_uint4(elem, step, (uint32_t)off, (uint32_t)(off >> 32)), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u1 = ((float)(r.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  const float u2 = ((float)(r.y >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

// LayerNorm statistics of one lane-strided row (eps 1e-6, diffloss.py:110,137), in place: u <- (u - mean) * rstd
template <int NCH>
__device__ __forceinline__ void dh_normalise(float (&u)[NCH * 4], int W) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NCH * 4; ++j) s += u[j];
  const float mean = warp_sum(s) / (float)W;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NCH * 4; ++j) {
    u[j] -= mean;
    q = fmaf(u[j], u[j], q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)W + 1e-6f);
#pragma unroll
  for (int j = 0; j < NCH * 4; ++j) u[j] *= rstd;
}

// input_proj of row r (diffloss.py:219) for this lane's columns, from the sample in shared memory
template <int NCH>
__device__ __forceinline__ void dh_input_proj(const DhParams& P, const float* xs_row, const float* WiT, int lane, float (&u)[NCH * 4]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int k = ch * 128 + lane * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(P.in_b + k));
    for (int c = 0; c < P.C; ++c) {
      const float x = xs_row[c];
      const float4 w = *reinterpret_cast<const float4*>(WiT + c * P.W + k);
      a.x = fmaf(x, w.x, a.x); a.y = fmaf(x, w.y, a.y); a.z = fmaf(x, w.z, a.z); a.w = fmaf(x, w.w, a.w);
    }
    u[ch * 4] = a.x; u[ch * 4 + 1] = a.y; u[ch * 4 + 2] = a.z; u[ch * 4 + 3] = a.w;
  }
}
// the same for ONE column n (the residual a phase-B finaliser needs in block 0): same order of operations as above
__device__ __forceinline__ float dh_input_proj_one(const DhParams& P, const float* xs_row, const float* WiT, int n) {
  float a = __ldg(P.in_b + n);
  for (int c = 0; c < P.C; ++c) a = fmaf(xs_row[c], WiT[c * P.W + n], a);
  return a;
}

// Modulation of one (row, LayerNorm) application, fetched ahead of the grid barrier in front of the phase that uses it (it
// depends on the step and the row, not on what the barrier orders): sc1 = 1 + scale, sh = shift, adaLN biases folded in.
template <int NCH> struct DhPre { float4 sc1[NCH], sh[NCH]; };
template <int NCH> struct DhAffine { float4 g[NCH], o[NCH]; };

template <int NCH>
__device__ __forceinline__ void dh_prefetch(const DhParams& P, const float* mrow, int sh_off, int sc_off, int lane, DhPre<NCH>& pre) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int k = ch * 128 + lane * 4;
    const float4 sh = __ldg(reinterpret_cast<const float4*>(mrow + sh_off + k)), shb = __ldg(reinterpret_cast<const float4*>(P.ada_b + sh_off + k));
    const float4 sc = __ldg(reinterpret_cast<const float4*>(mrow + sc_off + k)), scb = __ldg(reinterpret_cast<const float4*>(P.ada_b + sc_off + k));
    pre.sh[ch] = make_float4(sh.x + shb.x, sh.y + shb.y, sh.z + shb.z, sh.w + shb.w);
    pre.sc1[ch] = make_float4(1.0f + (sc.x + scb.x), 1.0f + (sc.y + scb.y), 1.0f + (sc.z + scb.z), 1.0f + (sc.w + scb.w));
  }
}
// in_ln's affine parameters (model constants: requested together with the activations, they arrive while the statistics are reduced)
template <int NCH>
__device__ __forceinline__ void dh_load_affine(const float* ln_w, const float* ln_b, int lane, DhAffine<NCH>& af) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const int k = ch * 128 + lane * 4;
    af.g[ch] = __ldg(reinterpret_cast<const float4*>(ln_w + k));
    af.o[ch] = __ldg(reinterpret_cast<const float4*>(ln_b + k));
  }
}
// u <- (u * g + o) * (1 + scale) + shift   (in_ln's affine, then modulate(): diffloss.py:55-56, 123)
template <int NCH>
__device__ __forceinline__ void dh_apply(const DhAffine<NCH>& af, const DhPre<NCH>& pre, float (&u)[NCH * 4]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    u[ch * 4 + 0] = fmaf(fmaf(u[ch * 4 + 0], af.g[ch].x, af.o[ch].x), pre.sc1[ch].x, pre.sh[ch].x);
    u[ch * 4 + 1] = fmaf(fmaf(u[ch * 4 + 1], af.g[ch].y, af.o[ch].y), pre.sc1[ch].y, pre.sh[ch].y);
    u[ch * 4 + 2] = fmaf(fmaf(u[ch * 4 + 2], af.g[ch].z, af.o[ch].z), pre.sc1[ch].z, pre.sh[ch].z);
    u[ch * 4 + 3] = fmaf(fmaf(u[ch * 4 + 3], af.g[ch].w, af.o[ch].w), pre.sc1[ch].w, pre.sh[ch].w);
  }
}
// the output layer's LayerNorm has no affine parameters (diffloss.py:137)
template <int NCH>
__device__ __forceinline__ void dh_apply(const DhPre<NCH>& pre, float (&u)[NCH * 4]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    u[ch * 4 + 0] = fmaf(u[ch * 4 + 0], pre.sc1[ch].x, pre.sh[ch].x);
    u[ch * 4 + 1] = fmaf(u[ch * 4 + 1], pre.sc1[ch].y, pre.sh[ch].y);
    u[ch * 4 + 2] = fmaf(u[ch * 4 + 2], pre.sc1[ch].z, pre.sh[ch].z);
    u[ch * 4 + 3] = fmaf(u[ch * 4 + 3], pre.sc1[ch].w, pre.sh[ch].w);
  }
}

template <int NCH>
__device__ __forceinline__ void dh_load_row_cg(const float* row, int lane, float (&u)[NCH * 4]) {
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
    const float4 t = dh_ldcg4(row + ch * 128 + lane * 4);
    u[ch * 4] = t.x; u[ch * 4 + 1] = t.y; u[ch * 4 + 2] = t.z; u[ch * 4 + 3] = t.w;
  }
}

// dot products of one register row with up to four consecutive shared-memory weight rows, reduced over the warp
template <int NCH>
__device__ __forceinline__ void dh_dot4(const float (&u)[NCH * 4], const float* w0, int W, int ncols, int lane, float (&acc)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = 0.f;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j >= ncols) continue;
      const float4 w = *reinterpret_cast<const float4*>(w0 + (size_t)j * W + ch * 128 + lane * 4);
      float a = acc[j];
      a = fmaf(u[ch * 4], w.x, a); a = fmaf(u[ch * 4 + 1], w.y, a); a = fmaf(u[ch * 4 + 2], w.z, a); a = fmaf(u[ch * 4 + 3], w.w, a);
      acc[j] = a;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = warp_sum(acc[j]);
}
__device__ __forceinline__ float dh_pick4(const float (&acc)[4], int lane) {
  float v = acc[0];
#pragma unroll
  for (int j = 1; j < 4; ++j) v = lane == j ? acc[j] : v;
  return v;
}

// Work split inside a column CTA: an item is (row, column group); a warp takes one item at a time, holds that ONE row in
// registers (lane-strided) and produces the group's columns.  With fewer rows than warps the CTA's columns are cut into
// kDhWarps / R groups so that every warp has something to do; from kDhWarps rows on, a warp owns whole rows.  (A first
// version kept four rows per warp in registers and gave every warp all the rows: at 255 registers ptxas could no longer
// batch the loads, and each of ~3000 instructions per warp and phase waited out its own L2 round trip -- 28 us per phase at
// 8 rows, ncu: 11 of 19 cycles per issue on the long scoreboard.)
struct DhSplit { int groups, gcols, items; };
__device__ __forceinline__ DhSplit dh_split(int R, int cpc) {
  DhSplit s;
  s.groups = R >= kDhWarps ? 1 : kDhWarps / R;
  s.gcols = (cpc + s.groups - 1) / s.groups;
  s.items = R * s.groups;
  return s;
}

template <int NCH>
__global__ void __launch_bounds__(kDhThreads, 1) dh_sample_kernel(const __grid_constant__ DhParams P) {
  extern __shared__ float dh_smem[];
  const int W = P.W, C = P.C, R = P.R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool col_cta = (int)blockIdx.x < P.G;
  // shared memory: [block weights: depth x (w1 | w2) x cpc x W  |  or the output-layer row: W] [WiT: C x W] [xs: R x C]
  const int wfloats = col_cta ? 2 * P.depth * P.cpc * W : W;
  float* wsm = dh_smem;
  float* WiT = dh_smem + wfloats;
  float* xs = WiT + C * W;

  // ---- one-time: this CTA's weight slices, the transposed input projection, the start sample
  if (col_cta) {
    const size_t row0 = (size_t)blockIdx.x * P.cpc;
    for (int b = 0; b < P.depth; ++b) {
      const float4* s1 = reinterpret_cast<const float4*>(P.blk[b].w1 + row0 * W);
      const float4* s2 = reinterpret_cast<const float4*>(P.blk[b].w2 + row0 * W);
      float4* d1 = reinterpret_cast<float4*>(wsm + (size_t)(2 * b) * P.cpc * W);
      float4* d2 = reinterpret_cast<float4*>(wsm + (size_t)(2 * b + 1) * P.cpc * W);
      for (int e = threadIdx.x; e < P.cpc * W / 4; e += kDhThreads) { d1[e] = __ldg(s1 + e); d2[e] = __ldg(s2 + e); }
    }
  } else {
    const float4* s = reinterpret_cast<const float4*>(P.final_w + (size_t)(blockIdx.x - P.G) * W);
    for (int e = threadIdx.x; e < W / 4; e += kDhThreads) reinterpret_cast<float4*>(wsm)[e] = __ldg(s + e);
  }
  for (int e = threadIdx.x; e < W * C; e += kDhThreads) WiT[(e % C) * W + e / C] = __ldg(P.in_w + e);
  const int half = P.use_cfg ? R / 2 : R;
  for (int e = threadIdx.x; e < R * C; e += kDhThreads) {
    const int r = e / C, c = e - r * C;
    xs[e] = P.noise ? __ldg(P.noise + e) : dh_normal(P.seed, P.offset, (uint32_t)P.steps, (uint32_t)((r % half) * C + c));
  }
  __syncthreads();

  const DhSplit sp = dh_split(R, P.cpc);
  // parameters of this warp's first item of the next LayerNorm phase, requested before the barrier in front of that phase
  DhPre<NCH> pre;
  const int r_first = col_cta ? warp / sp.groups : warp;
  const bool has_first = col_cta ? (warp < sp.items && (warp - r_first * sp.groups) * sp.gcols < P.cpc) : warp < R;
  auto prefetch_block = [&](int b, const float* ms) {
    if (col_cta && has_first) dh_prefetch<NCH>(P, ms + (size_t)r_first * P.n_mod, b * 3 * W, b * 3 * W + W, lane, pre);
  };
  prefetch_block(0, P.mod + (size_t)(P.steps - 1) * R * P.n_mod);
  unsigned bar_target = 0;
  for (int i = P.steps - 1; i >= 0; --i) {
    const float* mstep = P.mod + (size_t)i * R * P.n_mod;
    for (int b = 0; b < P.depth; ++b) {
      // ---------------- phase A: v = SiLU( modulate(LN(h)) . w1^T + b1 )                         (diffloss.py:122-125)
      if (col_cta) {
        const float* w1s = wsm + (size_t)(2 * b) * P.cpc * W;
        for (int it = warp; it < sp.items; it += kDhWarps) {
          const int r = it / sp.groups, q0 = (it - r * sp.groups) * sp.gcols, q1 = min(P.cpc, q0 + sp.gcols);
          if (q0 >= q1) continue;
          float u[NCH * 4];
          if (b == 0) dh_input_proj<NCH>(P, xs + (r % half) * C, WiT, lane, u);
          else dh_load_row_cg<NCH>(P.hbuf + (size_t)r * W, lane, u);
          DhAffine<NCH> af;
          dh_load_affine<NCH>(P.blk[b].ln_w, P.blk[b].ln_b, lane, af);
          if (it != warp) dh_prefetch<NCH>(P, mstep + (size_t)r * P.n_mod, b * 3 * W, b * 3 * W + W, lane, pre);
          dh_normalise<NCH>(u, W);
          dh_apply<NCH>(af, pre, u);
          for (int q = q0; q < q1; q += 4) {
            const int nc = min(4, q1 - q), n = blockIdx.x * P.cpc + q + lane;        // lanes [0, nc) finalise one column each
            const float bias = lane < nc ? __ldg(P.blk[b].b1 + n) : 0.f;
            float acc[4];
            dh_dot4<NCH>(u, w1s + (size_t)q * W, W, nc, lane, acc);
            if (lane < nc) P.vbuf[(size_t)r * W + n] = dh_silu(dh_pick4(acc, lane) + bias);
          }
        }
      }
      dh_grid_barrier(P.bar, bar_target, P.err);
      // ---------------- phase B: h <- h + gate * ( v . w2^T + b2 )                                (diffloss.py:125-126)
      if (col_cta) {
        const float* w2s = wsm + (size_t)(2 * b + 1) * P.cpc * W;
        for (int it = warp; it < sp.items; it += kDhWarps) {
          const int r = it / sp.groups, q0 = (it - r * sp.groups) * sp.gcols, q1 = min(P.cpc, q0 + sp.gcols);
          if (q0 >= q1) continue;
          float u[NCH * 4];
          dh_load_row_cg<NCH>(P.vbuf + (size_t)r * W, lane, u);
          for (int q = q0; q < q1; q += 4) {
            const int nc = min(4, q1 - q), n = blockIdx.x * P.cpc + q + lane;
            float gate = 0.f, res = 0.f, bias = 0.f;
            if (lane < nc) {
              gate = __ldg(mstep + (size_t)r * P.n_mod + b * 3 * W + 2 * W + n) + __ldg(P.ada_b + b * 3 * W + 2 * W + n);
              res = b == 0 ? dh_input_proj_one(P, xs + (r % half) * C, WiT, n) : dh_ldcg(P.hbuf + (size_t)r * W + n);
              bias = __ldg(P.blk[b].b2 + n);
            }
            float acc[4];
            dh_dot4<NCH>(u, w2s + (size_t)q * W, W, nc, lane, acc);
            if (lane < nc) P.hbuf[(size_t)r * W + n] = fmaf(gate, dh_pick4(acc, lane) + bias, res);
          }
        }
      }
      if (b + 1 < P.depth) prefetch_block(b + 1, mstep);
      else if (!col_cta && has_first) dh_prefetch<NCH>(P, mstep + (size_t)r_first * P.n_mod, P.depth * 3 * W, P.depth * 3 * W + W, lane, pre);
      dh_grid_barrier(P.bar, bar_target, P.err);
    }
    // ---------------- output layer: out = modulate(LN(h)) . final_w^T + final_b, one column per CTA   (diffloss.py:143-149)
    if (!col_cta) {
      const int col = blockIdx.x - P.G;
      for (int r = warp; r < R; r += kDhWarps) {
        float u[NCH * 4];
        dh_load_row_cg<NCH>(P.hbuf + (size_t)r * W, lane, u);
        dh_normalise<NCH>(u, W);
        if (r != warp) dh_prefetch<NCH>(P, mstep + (size_t)r * P.n_mod, P.depth * 3 * W, P.depth * 3 * W + W, lane, pre);
        dh_apply<NCH>(pre, u);
        float a = 0.f;
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const float4 w = *reinterpret_cast<const float4*>(wsm + ch * 128 + lane * 4);
          a = fmaf(u[ch * 4], w.x, a); a = fmaf(u[ch * 4 + 1], w.y, a); a = fmaf(u[ch * 4 + 2], w.z, a); a = fmaf(u[ch * 4 + 3], w.w, a);
        }
        a = warp_sum(a);
        if (lane == 0) P.obuf[r * 2 * C + col] = a + __ldg(P.final_b + col);
      }
    }
    if (i > 0) prefetch_block(0, P.mod + (size_t)(i - 1) * R * P.n_mod);
    dh_grid_barrier(P.bar, bar_target, P.err);
    // ---------------- posterior update, redundantly in every CTA   (gaussian_diffusion.py:283-291, 334-339, 238-241, 411-419)
    if (threadIdx.x < R * C) {
      const int r = threadIdx.x / C, c = threadIdx.x - r * C;
      float eps = dh_ldcg(P.obuf + r * 2 * C + c);
      const float vv = dh_ldcg(P.obuf + r * 2 * C + C + c);
      if (P.use_cfg) {                                           // diffloss.py:242-246
        const int rc = r % half;
        const float ce = dh_ldcg(P.obuf + rc * 2 * C + c), ue = dh_ldcg(P.obuf + (half + rc) * 2 * C + c);
        eps = __fadd_rn(ue, __fmul_rn(P.cfg_scale, __fsub_rn(ce, ue)));
      }
      const float a = __ldg(P.tab + i), bq = __ldg(P.tab + P.steps + i), c1 = __ldg(P.tab + 2 * P.steps + i), c2 = __ldg(P.tab + 3 * P.steps + i);
      const float min_log = __ldg(P.tab + 4 * P.steps + i), max_log = __ldg(P.tab + 5 * P.steps + i);
      const float x = xs[threadIdx.x];
      const float frac = __fdiv_rn(__fadd_rn(vv, 1.0f), 2.0f);
      const float logvar = __fadd_rn(__fmul_rn(frac, max_log), __fmul_rn(__fsub_rn(1.0f, frac), min_log));
      const float xstart = __fsub_rn(__fmul_rn(a, x), __fmul_rn(bq, eps));
      float xn = __fadd_rn(__fmul_rn(c1, xstart), __fmul_rn(c2, x));
      if (i != 0) {
        const float nz = P.step_noise ? __ldg(P.step_noise + ((size_t)i * R + r) * C + c) : dh_normal(P.seed, P.offset, (uint32_t)i, (uint32_t)threadIdx.x);
        xn = __fadd_rn(xn, __fmul_rn(__fmul_rn(expf(__fmul_rn(0.5f, logvar)), nz), P.temperature));
      }
      xs[threadIdx.x] = xn;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x < R * C) P.out[threadIdx.x] = xs[threadIdx.x];
  // leave the counters at zero for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(P.bar + 1, 1u);
    if (old == gridDim.x - 1) {
      P.bar[0] = 0;
      P.bar[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ small setup kernels
// sinusoidal timestep embedding (diffloss.py:71-92): out [n, 256] = (cos(t f_j), sin(t f_j)), f_j = exp(-ln(10000) j / 128)
__global__ void dh_timestep_embedding_kernel(const int* __restrict__ t, float* __restrict__ out, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  constexpr int half = kDhTimeFreq / 2;
  if (idx >= n * half) return;
  const int s = idx / half, j = idx - s * half;
  const float f = expf(-9.210340371976184f * (float)j / (float)half);
  const float a = (float)t[s] * f;
  out[(size_t)s * kDhTimeFreq + j] = cosf(a);
  out[(size_t)s * kDhTimeFreq + half + j] = sinf(a);
}
// out[m, n] = act( x[m, :] . w[n, :] + b[n] ), one warp per output; ACT: 0 none, 1 SiLU
template <int ACT>
__global__ void dh_linear_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                 float* __restrict__ out, int M, int N, int K) {
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= (long long)M * N) return;
  const int m = (int)(o / N), n = (int)(o - (long long)m * N);
  float a = 0.f;
  for (int k = lane; k < K; k += 32) a = fmaf(x[(size_t)m * K + k], w[(size_t)n * K + k], a);
  a = warp_sum(a);
  if (lane == 0) {
    a += b[n];
    out[o] = ACT == 1 ? dh_silu(a) : a;
  }
}
// ysilu[s * R + r, :] = SiLU(t_emb[s] + c_emb[r])   (y = t + c, diffloss.py:223; the SiLU in front of every adaLN Linear)
__global__ void dh_ysilu_kernel(const float* __restrict__ t_emb, const float* __restrict__ c_emb, float* __restrict__ out,
                                int steps, int R, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)steps * R * W) return;
  const int k = (int)(idx % W);
  const long long sr = idx / W;
  const int r = (int)(sr % R), s = (int)(sr / R);
  out[idx] = dh_silu(t_emb[(size_t)s * W + k] + c_emb[(size_t)r * W + k]);
}


// ------------------------------------------------------------------------------------------------ modulation GEMM
// C[M, N] = A[M, K] . B[N, K]^T in fp32 on the CUDA cores (no TF32: the head's parity contract is fp32): every adaLN
// modulation of a token's chain, M = steps * rows, N = (3 depth + 2) W, K = W.  128 x 128 tile, 16-deep k-slices double
// buffered in shared memory (stored k-major so a thread's 8 + 8 operands are four conflict-free 128-bit loads), 8 x 8
// outputs per thread as two 4-wide column groups 64 apart.  M is ragged (clamped loads, guarded stores); N % 128 == 0,
// K % 16 == 0.
constexpr int kSgBM = 128, kSgBN = 128, kSgBK = 16, kSgThreads = 256;

__global__ void __launch_bounds__(kSgThreads) dh_sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ Cm, int M, int N, int K, int ldc) {
  __shared__ __align__(16) float As[2][kSgBK][kSgBM];
  __shared__ __align__(16) float Bs[2][kSgBK][kSgBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kSgBM, n0 = blockIdx.x * kSgBN;
  // loader: thread -> (row = tid % 128, k-quad = tid / 128 and + 2): consecutive threads write consecutive shared-memory words
  const int lrow = tid & 127, lq = tid >> 7;
  const float* ap = A + (size_t)min(m0 + lrow, M - 1) * K + lq * 4;
  const float* bp = B + (size_t)(n0 + lrow) * K + lq * 4;
  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
    ra[0] = __ldg(reinterpret_cast<const float4*>(ap + k0)); ra[1] = __ldg(reinterpret_cast<const float4*>(ap + k0 + 8));
    rb[0] = __ldg(reinterpret_cast<const float4*>(bp + k0)); rb[1] = __ldg(reinterpret_cast<const float4*>(bp + k0 + 8));
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int kq = (lq + 2 * j) * 4;
      As[buf][kq][lrow] = ra[j].x; As[buf][kq + 1][lrow] = ra[j].y; As[buf][kq + 2][lrow] = ra[j].z; As[buf][kq + 3][lrow] = ra[j].w;
      Bs[buf][kq][lrow] = rb[j].x; Bs[buf][kq + 1][lrow] = rb[j].y; Bs[buf][kq + 2][lrow] = rb[j].z; Bs[buf][kq + 3][lrow] = rb[j].w;
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = K / kSgBK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * kSgBK);
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]), b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
    if (m >= M) continue;
    float* c = Cm + (size_t)m * ldc + n0;
    *reinterpret_cast<float4*>(c + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(c + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

}  // namespace vlg

// ================================================================================================== C ABI
using namespace vlg;

struct vlg_diff {
  vlg_diff_geometry g;
  vlg_diff_weights w;
  std::vector<vlg_diff_block> blocks;
  int max_rows = 0, cpc = 0, G = 0, nch = 0, n_mod = 0, smem = 0, device = 0;
  uint8_t* ws = nullptr;
  // workspace regions
  float *t_emb = nullptr, *tab = nullptr, *ada_w = nullptr, *ada_b = nullptr, *c_emb = nullptr, *ysilu = nullptr, *mod = nullptr;
  float *hbuf = nullptr, *vbuf = nullptr, *obuf = nullptr;
  unsigned* bar = nullptr;
  unsigned* err_host = nullptr;
  unsigned* err_dev = nullptr;
};

namespace {

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

struct DhLayout { size_t t_emb, tab, ada_w, ada_b, c_emb, ysilu, mod, hbuf, vbuf, obuf, bar, tfreq, tmid, tmap, total; };

DhLayout dh_layout(const vlg_diff_geometry& g, int max_rows) {
  DhLayout L{};
  const size_t W = g.W, n_mod = (size_t)(3 * g.depth + 2) * g.W, S = g.steps, R = max_rows;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes); return at; };
  L.t_emb = take(S * W * 4);
  L.tab = take(6 * S * 4);
  L.ada_w = take(n_mod * W * 4);      // the adaLN weight matrices of all blocks and the final layer, one under the other
  L.ada_b = take(n_mod * 4);
  L.c_emb = take(R * W * 4);
  L.ysilu = take(S * R * W * 4);
  L.mod = take(S * R * n_mod * 4);
  L.hbuf = take(R * W * 4);
  L.vbuf = take(R * W * 4);
  L.obuf = take(R * 2 * g.C * 4);
  L.bar = take(256);
  L.tfreq = take(S * kDhTimeFreq * 4);     // create-time scratch: sinusoid, hidden of the timestep MLP, timestep map
  L.tmid = take(S * W * 4);
  L.tmap = take(S * 4);
  L.total = o;
  return L;
}

// columns per CTA: the fewest (most CTAs) that leaves room for the 2C output-layer CTAs and fits shared memory
int dh_plan(const vlg_diff_geometry& g, int sms, int* cpc_out, int* G_out, int* smem_out) {
  for (int cpc = 1; cpc <= g.W; ++cpc) {
    if (g.W % cpc) continue;
    const int G = g.W / cpc;
    if (G + 2 * g.C > sms) continue;
    const size_t fl = (size_t)2 * g.depth * cpc * g.W + (size_t)g.C * g.W + (size_t)kDhMaxRows * g.C + 64;
    if (fl * 4 > 227 * 1024) return VLG_ERR_ARG;
    *cpc_out = cpc; *G_out = G; *smem_out = (int)(fl * 4);
    return VLG_OK;
  }
  return VLG_ERR_ARG;
}

int dh_check_geometry(const vlg_diff_geometry* g, int max_rows) {
  VLG_REQUIRE(g != nullptr, "diff: null geometry");
  VLG_REQUIRE(g->C >= 1 && g->C <= kDhMaxC, "diff: target_channels %d not in [1, %d]", g->C, kDhMaxC);
  VLG_REQUIRE(g->W % 128 == 0 && (g->W / 128 == 1 || g->W / 128 == 2 || g->W / 128 == 4 || g->W / 128 == 6 || g->W / 128 == 8),
              "diff: width %d unsupported (128, 256, 512, 768 or 1024)", g->W);
  VLG_REQUIRE(g->Z >= 1, "diff: z_channels %d", g->Z);
  VLG_REQUIRE(g->depth >= 1 && g->depth <= kDhMaxDepth, "diff: depth %d not in [1, %d]", g->depth, kDhMaxDepth);
  VLG_REQUIRE(g->steps >= 1 && g->steps <= 1000, "diff: steps %d not in [1, 1000]", g->steps);
  VLG_REQUIRE(max_rows >= 1 && max_rows <= kDhMaxRows, "diff: max_rows %d not in [1, %d]", max_rows, kDhMaxRows);
  return VLG_OK;
}

template <int NCH>
int dh_launch(const vlg_diff* d, const DhParams& P, cudaStream_t st) {
  auto kern = dh_sample_kernel<NCH>;
  static PerDeviceOnce once;
  VLG_CUDA(once([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); }));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(P.G + 2 * P.C);
  cfg.blockDim = dim3(kDhThreads);
  cfg.dynamicSmemBytes = (size_t)d->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs resident at once, or the launch fails: the barriers need it
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VLG_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

int dh_err_check(vlg_diff* d) {
  if (d->err_host == nullptr || *reinterpret_cast<volatile unsigned*>(d->err_host) == 0) return VLG_OK;
  cudaDeviceSynchronize();
  *d->err_host = 0;
  cudaMemset(d->bar, 0, 256);
  set_error("a diffusion-head launch timed out in its grid barrier (its CTAs were not co-resident); outputs since the previous "
            "successful call are invalid");
  return VLG_ERR_STATE;
}

}  // namespace

extern "C" {

size_t vlg_diff_workspace_bytes(const vlg_diff_geometry* g, int max_rows) {
  if (dh_check_geometry(g, max_rows) != VLG_OK) return 0;
  return dh_layout(*g, max_rows).total;
}

int vlg_diff_create(const vlg_diff_geometry* g, const vlg_diff_weights* w, const int32_t* timestep_map_host,
                    const float* tables_host, int max_rows, void* workspace_dev, size_t workspace_bytes, vlg_diff** out) {
  VLG_REQUIRE(out != nullptr && w != nullptr && timestep_map_host != nullptr && tables_host != nullptr, "diff_create: null argument");
  *out = nullptr;
  int rc = dh_check_geometry(g, max_rows);
  if (rc) return rc;
  int dev = 0;
  VLG_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VLG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("diff_create: device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor);
    return VLG_ERR_UNSUPPORTED;
  }
  const DhLayout L = dh_layout(*g, max_rows);
  VLG_REQUIRE(workspace_dev != nullptr && workspace_bytes >= L.total && (reinterpret_cast<uintptr_t>(workspace_dev) & 255) == 0,
              "diff_create: workspace too small or misaligned (%zu < %zu)", workspace_bytes, L.total);
  VLG_REQUIRE(w->blocks != nullptr && w->time_w0 && w->time_b0 && w->time_w1 && w->time_b1 && w->cond_w && w->cond_b && w->in_w &&
                  w->in_b && w->final_ada_w && w->final_ada_b && w->final_w && w->final_b, "diff_create: null weight pointer");
  vlg_diff* d = new vlg_diff();
  d->g = *g;
  d->w = *w;
  d->blocks.assign(w->blocks, w->blocks + g->depth);
  d->w.blocks = d->blocks.data();
  d->max_rows = max_rows;
  d->device = dev;
  d->nch = g->W / 128;
  d->n_mod = (3 * g->depth + 2) * g->W;
  if (dh_plan(*g, num_sms(), &d->cpc, &d->G, &d->smem) != VLG_OK) {
    set_error("diff_create: %d blocks of width %d do not fit the shared memory of %d SMs (fp32 weights stay resident for the whole chain)",
              g->depth, g->W, num_sms());
    delete d;
    return VLG_ERR_ARG;
  }
  d->ws = static_cast<uint8_t*>(workspace_dev);
  auto at = [&](size_t off) { return reinterpret_cast<float*>(d->ws + off); };
  d->t_emb = at(L.t_emb); d->tab = at(L.tab); d->ada_w = at(L.ada_w); d->ada_b = at(L.ada_b); d->c_emb = at(L.c_emb); d->ysilu = at(L.ysilu); d->mod = at(L.mod);
  d->hbuf = at(L.hbuf); d->vbuf = at(L.vbuf); d->obuf = at(L.obuf);
  d->bar = reinterpret_cast<unsigned*>(d->ws + L.bar);
  auto fail = [&](int code) { vlg_diff_destroy(d); return code; };
  if (cudaHostAlloc(reinterpret_cast<void**>(&d->err_host), sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) return fail(VLG_ERR_CUDA);
  *d->err_host = 0;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d->err_dev), d->err_host, 0) != cudaSuccess) return fail(VLG_ERR_CUDA);
  cudaStream_t st = nullptr;
  const int S = g->steps, W = g->W;
#define DH_TRY(expr) do { if ((expr) != cudaSuccess) { cuda_fail(cudaGetLastError(), #expr, __FILE__, __LINE__); return fail(VLG_ERR_CUDA); } } while (0)
  DH_TRY(cudaMemsetAsync(d->bar, 0, 256, st));
  DH_TRY(cudaMemcpyAsync(d->tab, tables_host, sizeof(float) * 6 * S, cudaMemcpyHostToDevice, st));
  int* tmap = reinterpret_cast<int*>(d->ws + L.tmap);
  DH_TRY(cudaMemcpyAsync(tmap, timestep_map_host, sizeof(int) * S, cudaMemcpyHostToDevice, st));
  // adaLN weights and biases, concatenated in the column order of the modulation buffer (one GEMM per token instead of depth + 1)
  for (int b = 0; b < g->depth; ++b) {
    if (!d->blocks[b].ada_w || !d->blocks[b].ada_b) { set_error("diff_create: null adaLN weight pointer in block %d", b); return fail(VLG_ERR_ARG); }
    DH_TRY(cudaMemcpyAsync(d->ada_w + (size_t)b * 3 * W * W, d->blocks[b].ada_w, sizeof(float) * 3 * W * W, cudaMemcpyDeviceToDevice, st));
    DH_TRY(cudaMemcpyAsync(d->ada_b + (size_t)b * 3 * W, d->blocks[b].ada_b, sizeof(float) * 3 * W, cudaMemcpyDeviceToDevice, st));
  }
  DH_TRY(cudaMemcpyAsync(d->ada_w + (size_t)g->depth * 3 * W * W, w->final_ada_w, sizeof(float) * 2 * W * W, cudaMemcpyDeviceToDevice, st));
  DH_TRY(cudaMemcpyAsync(d->ada_b + (size_t)g->depth * 3 * W, w->final_ada_b, sizeof(float) * 2 * W, cudaMemcpyDeviceToDevice, st));
  // timestep-embedding table of the model: t_emb[s] = Linear(SiLU(Linear(sinusoid(timestep_map[s]))))   (diffloss.py:94-97)
  float* tfreq = at(L.tfreq);
  float* tmid = at(L.tmid);
  dh_timestep_embedding_kernel<<<(S * (kDhTimeFreq / 2) + 255) / 256, 256, 0, st>>>(tmap, tfreq, S);
  count_launches(1);
  dh_linear_kernel<1><<<(S * W + 7) / 8, 256, 0, st>>>(tfreq, w->time_w0, w->time_b0, tmid, S, W, kDhTimeFreq);
  count_launches(1);
  dh_linear_kernel<0><<<(S * W + 7) / 8, 256, 0, st>>>(tmid, w->time_w1, w->time_b1, d->t_emb, S, W, W);
  count_launches(1);
  DH_TRY(cudaGetLastError());
  DH_TRY(cudaStreamSynchronize(st));
#undef DH_TRY
  *out = d;
  return VLG_OK;
}

void vlg_diff_destroy(vlg_diff* d) {
  if (d == nullptr) return;
  if (d->err_host) cudaFreeHost(d->err_host);
  delete d;
}

int vlg_diff_check(vlg_diff* d, void* stream) {
  VLG_REQUIRE(d != nullptr, "diff_check: null handle");
  VLG_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return dh_err_check(d);
}

int vlg_diff_sample(vlg_diff* d, const float* z_dev, int rows, float temperature, float cfg_scale, const float* noise_dev,
                    const float* step_noise_dev, uint64_t seed, uint64_t offset, float* out_dev, void* stream) {
  VLG_REQUIRE(d != nullptr && z_dev != nullptr && out_dev != nullptr, "diff_sample: null argument");
  VLG_REQUIRE(rows >= 1 && rows <= d->max_rows, "diff_sample: rows %d not in [1, %d]", rows, d->max_rows);
  const bool use_cfg = cfg_scale != 1.0f;
  VLG_REQUIRE(!use_cfg || rows % 2 == 0, "diff_sample: guidance needs an even row count (conditional rows, then unconditional), got %d", rows);
  int rc = dh_err_check(d);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const vlg_diff_geometry& g = d->g;
  const int W = g.W, S = g.steps;
  // conditioning of this token and every modulation of the chain
  dh_linear_kernel<0><<<(rows * W + 7) / 8, 256, 0, st>>>(z_dev, d->w.cond_w, d->w.cond_b, d->c_emb, rows, W, g.Z);
  count_launches(1);
  {
    const long long n = (long long)S * rows * W;
    dh_ysilu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d->t_emb, d->c_emb, d->ysilu, S, rows, W);
    count_launches(1);
  }
  VLG_CUDA(cudaGetLastError());
  dh_sgemm_kernel<<<dim3(d->n_mod / kSgBN, (S * rows + kSgBM - 1) / kSgBM), kSgThreads, 0, st>>>(d->ysilu, d->ada_w, d->mod, S * rows, d->n_mod, W,
                                                                                                 d->n_mod);
  count_launches(1);
  VLG_CUDA(cudaGetLastError());
  DhParams P{};
  P.R = rows; P.C = g.C; P.W = W; P.depth = g.depth; P.steps = S;
  P.cpc = d->cpc; P.G = d->G; P.n_mod = d->n_mod;
  P.temperature = temperature; P.cfg_scale = cfg_scale; P.use_cfg = use_cfg ? 1 : 0;
  P.in_w = d->w.in_w; P.in_b = d->w.in_b;
  for (int b = 0; b < g.depth; ++b) {
    const vlg_diff_block& s = d->blocks[b];
    VLG_REQUIRE(s.ln_w && s.ln_b && s.w1 && s.b1 && s.w2 && s.b2 && s.ada_w && s.ada_b, "diff_sample: null weight pointer in block %d", b);
    P.blk[b] = DhBlock{s.ln_w, s.ln_b, s.w1, s.b1, s.w2, s.b2};
  }
  P.final_w = d->w.final_w; P.final_b = d->w.final_b;
  P.ada_b = d->ada_b; P.mod = d->mod; P.tab = d->tab;
  P.noise = noise_dev; P.step_noise = step_noise_dev;
  P.seed = seed; P.offset = offset;
  P.hbuf = d->hbuf; P.vbuf = d->vbuf; P.obuf = d->obuf; P.out = out_dev;
  P.bar = d->bar; P.err = d->err_dev;
  switch (d->nch) {
    case 1: return dh_launch<1>(d, P, st);
    case 2: return dh_launch<2>(d, P, st);
    case 4: return dh_launch<4>(d, P, st);
    case 6: return dh_launch<6>(d, P, st);
    case 8: return dh_launch<8>(d, P, st);
  }
  set_error("diff_sample: width %d unsupported", W);
  return VLG_ERR_ARG;
}

}  // extern "C"
