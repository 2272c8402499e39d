// GEMM epilogues shared by the tcgen05 (bf16) and the exact fp32 GEMM kernels.
// Each epilogue consumes 8 consecutive output columns of one row and reproduces the rounding points
// of the reference op sequence it fuses: the nn.Linear result is rounded to the storage
// type T first (`rnd<T>`), then the fused elementwise ops run in fp32 and round once more.
#pragma once
#include "common.cuh"

namespace vlg {

enum Epi : int {
  EPI_STORE = 0,      // out[T]        = T(acc)
  EPI_STORE_F32 = 1,  // out[float]    = float(T(acc))            logits = output(h).float()      gpt.py:371
  EPI_RESID = 2,      // out[T]        = T(resid + T(acc))        h + wo(..), h + w2(..)          gpt.py:257-258
  EPI_SWIGLU = 3,     // out[T][n/2]   = T(T(silu(T(a))) * T(b))  silu(w1 x) * w3 x, w1/w3 rows interleaved  gpt.py:167
  EPI_QKV_ROPE = 4,   // RoPE(q,k) -> q buffer / KV cache write   gpt.py:215-227, 423-433, 177-185
  EPI_GELU = 5,       // out[T]        = T(gelu_tanh(T(acc)))     MLP.fc1 + act                   gpt.py:118-131
};

struct EpiParams {
  void* out;            // [rows, ld_out]
  int ld_out;
  const void* resid;    // EPI_RESID: [rows, ld_out]
  // EPI_QKV_ROPE
  void* q_out;          // [rows, dim]
  void* k_cache;        // this layer: [rows, H, S_max, D]
  void* v_cache;
  const float* rope;    // [rope_rows, D/2, 2]
  const int* pos_ptr;   // device scalar with the cache slot, or nullptr -> pos_imm
  int pos_imm;
  int S_max, H, D, dim;
  int q_per_row;        // batched prefill: GEMM row r is query (r % q_per_row) of cache row (r / q_per_row) at position
                        // pos_imm + r % q_per_row; 0 = decode (row r is cache row r at the scalar position)
  // Fused RMSNorm prologue (bf16 tensor-core GEMM only; norm_flag == nullptr -> the A operand is used as given).
  // The first CTAs of the grid normalise `norm_src` into the A buffer (the tensor map's base) and publish a counter;
  // every CTA's producer waits on it before loading A.  Same arithmetic as rmsnorm_kernel (gpt.py:143-148).
  const void* norm_src;     // [rows, K] un-normalised hidden state
  const void* norm_weight;  // [K]
  void* norm_dst;           // [rows, K] = the buffer the A tensor map reads
  unsigned* norm_flag;      // zero before the launch; CTAs that normalised rows add 1 each
  float norm_eps;
};

template <typename T> __device__ __forceinline__ void store8(T* dst, const float* v);
template <> __device__ __forceinline__ void store8<float>(float* dst, const float* v) {
  reinterpret_cast<float4*>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
  u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
template <typename T> __device__ __forceinline__ void load8(const T* src, float* v);
template <> __device__ __forceinline__ void load8<float>(const float* src, float* v) {
  float4 a = reinterpret_cast<const float4*>(src)[0], b = reinterpret_cast<const float4*>(src)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* src, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(src);
  v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
  v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
}

__device__ __forceinline__ float silu_f(float x) { return __fdiv_rn(x, 1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;   // ATen's GeluCUDAKernel form
  float inner = kBeta * (x + kKappa * (x * x * x));
  return 0.5f * x * (1.0f + tanhf(inner));
}

// r: output row; n0: first of 8 consecutive GEMM output columns; v: fp32 accumulators
template <int EPI, typename T>
__device__ __forceinline__ void epi8(int r, int n0, const float* v, const EpiParams& P) {
  if constexpr (EPI == EPI_STORE) {
    store8<T>(static_cast<T*>(P.out) + (size_t)r * P.ld_out + n0, v);
  } else if constexpr (EPI == EPI_STORE_F32) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = rnd<T>(v[i]);
    store8<float>(static_cast<float*>(P.out) + (size_t)r * P.ld_out + n0, o);
  } else if constexpr (EPI == EPI_RESID) {
    float o[8];
    load8<T>(static_cast<const T*>(P.resid) + (size_t)r * P.ld_out + n0, o);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __fadd_rn(o[i], rnd<T>(v[i]));
    store8<T>(static_cast<T*>(P.out) + (size_t)r * P.ld_out + n0, o);
  } else if constexpr (EPI == EPI_GELU) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = gelu_tanh_f(rnd<T>(v[i]));
    store8<T>(static_cast<T*>(P.out) + (size_t)r * P.ld_out + n0, o);
  } else if constexpr (EPI == EPI_SWIGLU) {
    T* dst = static_cast<T*>(P.out) + (size_t)r * P.ld_out + (n0 >> 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = rnd<T>(v[2 * i]), b = rnd<T>(v[2 * i + 1]);
      float s = rnd<T>(silu_f(a));
      dst[i] = Num<T>::from_f(__fmul_rn(s, b));
    }
  } else if constexpr (EPI == EPI_QKV_ROPE) {
    int pos, rc = r;
    if (P.q_per_row > 0) {
      rc = r / P.q_per_row;
      pos = P.pos_imm + (r - rc * P.q_per_row);
    } else {
      pos = P.pos_ptr ? *P.pos_ptr : P.pos_imm;
    }
    const int sec = n0 / P.dim, c = n0 - sec * P.dim;
    const int head = c / P.D, dd = c - head * P.D;
    float o[8];
    if (sec < 2) {
      const float2* tab = reinterpret_cast<const float2*>(P.rope) + (size_t)pos * (P.D >> 1) + (dd >> 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 cs = tab[i];
        float x0 = rnd<T>(v[2 * i]), x1 = rnd<T>(v[2 * i + 1]);
        // every product rounded on its own, as the eager mul/sub kernels do (no FMA contraction)
        o[2 * i] = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
        o[2 * i + 1] = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = v[i];
    }
    if (sec == 0) {
      store8<T>(static_cast<T*>(P.q_out) + (size_t)r * P.dim + c, o);
    } else {
      T* cache = static_cast<T*>(sec == 1 ? P.k_cache : P.v_cache);
      store8<T>(cache + (((size_t)rc * P.H + head) * P.S_max + pos) * P.D + dd, o);
    }
  }
}

}  // namespace vlg
