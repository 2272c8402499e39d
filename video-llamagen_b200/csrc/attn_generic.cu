// Single-query attention over the KV-cache prefix [0, pos] -- generic version (fp32 or bf16 storage).
//
// Replaces gpt.py:230-237 at seqlen 1: the reference copies both caches (repeat_interleave) and runs
// math-SDPA over all S_max slots with a boolean mask; here only the valid prefix is read, once.
// Numerics follow torch's math SDPA: fp32 throughout, q and k each pre-scaled by
// sqrt(1/sqrt(D)), masked keys excluded, one rounding of the output to the storage type.
//
// Mapping: one CTA per (row, head, split); 8 lanes share one cache position (D/8 contiguous elements
// each, 16-byte loads for bf16), so a warp touches 4 consecutive positions = one contiguous 512 B (bf16)
// span per load instruction.  Online softmax per lane group, merged across groups / warps at the end;
// sequence splits write (m, l, o) partials that a small combine kernel folds.
#include "common.cuh"
#include "kernels.h"

namespace vlg {

constexpr int kAttnThreads = 128;
constexpr int kMaxSplits = 32;

template <typename T, int D> struct KvVec;   // D/8 contiguous elements of one cache row -> fp32
template <> struct KvVec<__nv_bfloat16, 64> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) { load8<__nv_bfloat16>(p, v); }
};
template <> struct KvVec<float, 64> {
  static __device__ __forceinline__ void load(const float* p, float* v) { load8<float>(p, v); }
};

struct Partial { float m, l; };

template <typename T, int D>
__global__ void __launch_bounds__(kAttnThreads)
attn_decode_generic_kernel(const T* __restrict__ q, const T* __restrict__ kc, const T* __restrict__ vc,
                           T* __restrict__ out, float* __restrict__ part, int H, int S_max, const int* pos_ptr,
                           int pos_imm, const uint8_t* __restrict__ mask, int T_prefix, int n_splits, int q_per_row) {
  static_assert(D == 64, "head_dim 64");
  constexpr int E = D / 8;  // elements per lane
  __shared__ float s_m[kAttnThreads / 32], s_l[kAttnThreads / 32], s_o[kAttnThreads / 32][D];

  const int rh = blockIdx.x, split = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  // decode: query row == cache row, one position for all.  Batched prefill (q_per_row > 0): query row qr is position
  // pos_imm + qr % q_per_row of cache row qr / q_per_row.
  const int qr = rh / H, head = rh - qr * H;
  const int row = q_per_row > 0 ? qr / q_per_row : qr;
  const int pos = q_per_row > 0 ? pos_imm + (qr - row * q_per_row) : (pos_ptr ? *pos_ptr : pos_imm);
  const size_t kv_rh = (size_t)row * H + head;
  const int n = pos + 1;
  int chunk = (n + n_splits - 1) / n_splits;
  chunk = (chunk + 15) & ~15;
  const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 3, c = lane & 7;
  const float scale = sqrtf(1.0f / sqrtf((float)D));

  float qf[E];
  KvVec<T, D>::load(q + (size_t)rh * D + c * E, qf);
#pragma unroll
  for (int i = 0; i < E; ++i) qf[i] = __fmul_rn(qf[i], scale);

  const T* kbase = kc + kv_rh * S_max * D + c * E;
  const T* vbase = vc + kv_rh * S_max * D + c * E;
  const uint8_t* mrow = mask ? mask + (size_t)row * T_prefix : nullptr;

  float m = -INFINITY, l = 0.f, acc[E];
#pragma unroll
  for (int i = 0; i < E; ++i) acc[i] = 0.f;

  for (int sb = s_begin + warp * 4; sb < s_end; sb += (kAttnThreads / 32) * 4) {
    const int s = sb + g;
    const bool in = s < s_end;
    const int sc = in ? s : s_end - 1;
    float kf[E], vf[E];
    KvVec<T, D>::load(kbase + (size_t)sc * D, kf);
    KvVec<T, D>::load(vbase + (size_t)sc * D, vf);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i) dot = __fmaf_rn(qf[i], __fmul_rn(kf[i], scale), dot);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    // key visibility: causal prefix + emb_masks on the conditioning slots, diagonal always on (generate.py:156-165)
    bool vis = in && (s == pos || s >= T_prefix || mrow == nullptr || mrow[s] != 0);
    if (vis) {
      float m_new = fmaxf(m, dot);
      float corr = expf(m - m_new);  // m = -inf on first hit -> 0
      float p = expf(dot - m_new);
      l = __fmaf_rn(l, corr, p);
#pragma unroll
      for (int i = 0; i < E; ++i) acc[i] = __fmaf_rn(acc[i], corr, __fmul_rn(p, vf[i]));
      m = m_new;
    }
  }

  // merge the 4 position groups of the warp (lanes with equal c)
#pragma unroll
  for (int off = 8; off <= 16; off <<= 1) {
    float m2 = __shfl_xor_sync(0xffffffffu, m, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
    float mm = fmaxf(m, m2);
    float a = (m == -INFINITY) ? 0.f : expf(m - mm), b = (m2 == -INFINITY) ? 0.f : expf(m2 - mm);
    l = l * a + l2 * b;
#pragma unroll
    for (int i = 0; i < E; ++i) {
      float o2 = __shfl_xor_sync(0xffffffffu, acc[i], off);
      acc[i] = acc[i] * a + o2 * b;
    }
    m = mm;
  }
  if (lane < 8) {
    if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
#pragma unroll
    for (int i = 0; i < E; ++i) s_o[warp][c * E + i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < D) {
    const int d = threadIdx.x;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnThreads / 32; ++w) mm = fmaxf(mm, s_m[w]);
    float lt = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnThreads / 32; ++w) {
      float a = (s_m[w] == -INFINITY) ? 0.f : expf(s_m[w] - mm);
      lt += s_l[w] * a;
      o += s_o[w][d] * a;
    }
    if (n_splits == 1) {
      out[(size_t)rh * D + d] = Num<T>::from_f(__fdiv_rn(o, lt));
    } else {
      float* pp = part + ((size_t)rh * n_splits + split) * (D + 2);
      pp[2 + d] = o;
      if (d == 0) { pp[0] = mm; pp[1] = lt; }
    }
  }
}

template <typename T, int D>
__global__ void attn_combine_kernel(const float* __restrict__ part, T* __restrict__ out, int n_splits) {
  const int rh = blockIdx.x, d = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const float* pp = part + (size_t)rh * n_splits * (D + 2);
  float mm = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mm = fmaxf(mm, pp[s * (D + 2)]);
  float lt = 0.f, o = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    float ms = pp[s * (D + 2)];
    float a = (ms == -INFINITY) ? 0.f : expf(ms - mm);
    lt += pp[s * (D + 2) + 1] * a;
    o += pp[s * (D + 2) + 2 + d] * a;
  }
  out[(size_t)rh * D + d] = Num<T>::from_f(__fdiv_rn(o, lt));
}

size_t attn_scratch_bytes(int rows, int H, int D, int S_max) {
  (void)S_max;
  return (size_t)rows * H * kMaxSplits * (D + 2) * sizeof(float);
}

static int pick_splits(int rh, int max_pos) {
  // enough CTAs for ~2 waves; never more splits than 64-position chunks
  int want = (2 * num_sms() + rh - 1) / rh;
  int by_len = (max_pos + 1 + 63) / 64;
  int s = want < by_len ? want : by_len;
  if (s < 1) s = 1;
  if (s > kMaxSplits) s = kMaxSplits;
  return s;
}

template <typename T>
static int launch_typed(const AttnArgs& a, cudaStream_t st) {
  const int rh = a.rows * a.H * (a.q_per_row > 0 ? a.q_per_row : 1);
  int n_splits = a.q_per_row > 0 ? 1 : pick_splits(rh, a.max_pos);
  if (n_splits > 1) {
    VLG_REQUIRE(a.scratch && a.scratch_bytes >= (size_t)rh * n_splits * (a.D + 2) * sizeof(float),
                "attn: scratch too small for %d splits", n_splits);
  }
  VLG_CUDA(launch_kernel(attn_decode_generic_kernel<T, 64>, dim3(rh, n_splits), dim3(kAttnThreads), 0, st, (const T*)a.q,
                         (const T*)a.k_cache, (const T*)a.v_cache, (T*)a.out, a.scratch, a.H, a.S_max, a.pos_ptr,
                         a.pos_imm, a.prefix_mask, a.T_prefix, n_splits, a.q_per_row));
  VLG_LAUNCH_CHECK();
  if (n_splits > 1) {
    VLG_CUDA(launch_kernel(attn_combine_kernel<T, 64>, dim3(rh), dim3(64), 0, st, (const float*)a.scratch, (T*)a.out, n_splits));
    VLG_LAUNCH_CHECK();
  }
  return VLG_OK;
}

int launch_attn_generic(int dtype, const AttnArgs& a, cudaStream_t st) {
  VLG_REQUIRE(a.D == 64, "attn: head_dim %d not supported (64)", a.D);
  VLG_REQUIRE(a.rows > 0 && a.H > 0 && a.S_max > 0, "attn: bad shape");
  if (dtype == VLG_BF16) return launch_typed<__nv_bfloat16>(a, st);
  if (dtype == VLG_F32) return launch_typed<float>(a, st);
  set_error("attn: bad dtype %d", dtype);
  return VLG_ERR_ARG;
}

}  // namespace vlg
