// Layer-tail kernel: unit geometry and memory helpers (layer_tail.cu).
#pragma once
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vlg {

struct LtGemm {
  int N, K;        // out features, in features
  int BN;          // column tile (32, 64 or 128)
  int S;           // K splits
  int n_tiles;     // N / BN
  int kb_per;      // 64-wide k-blocks per split (the last split may own fewer)
};

struct LtUnit { int valid, n0, split, kb0, nkb; };

__device__ __forceinline__ LtUnit lt_unit(const LtGemm& G, int cta) {
  LtUnit u;
  u.valid = cta < G.n_tiles * G.S;
  const int tile = cta % G.n_tiles;      // neighbouring CTAs share the activation slice and differ in the weight tile
  u.split = cta / G.n_tiles;
  u.n0 = tile * G.BN;
  u.kb0 = u.split * G.kb_per;
  const int nkb = G.K / 64;
  u.nkb = min(G.kb_per, nkb - u.kb0);
  return u;
}

__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.cg.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
// 16 bf16 (32 bytes = one full L2 sector) in one store
__device__ __forceinline__ void stg256_bf16x16(__nv_bfloat16* p, const float* o) {
  asm volatile("st.global.cg.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(pack_bf16(o[0], o[1])), "r"(pack_bf16(o[2], o[3])),
               "r"(pack_bf16(o[4], o[5])), "r"(pack_bf16(o[6], o[7])), "r"(pack_bf16(o[8], o[9])), "r"(pack_bf16(o[10], o[11])),
               "r"(pack_bf16(o[12], o[13])), "r"(pack_bf16(o[14], o[15])) : "memory");
}
// L2-coherent loads: everything a row phase reads was written by other SMs earlier in this same kernel
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
__device__ __forceinline__ uint4 ldg128_cg(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// sum of the S <= kLtMaxSplits partials of 8 consecutive columns, fixed order; all loads in flight before the first add
constexpr int kLtMaxSplits = 8;
__device__ __forceinline__ void lt_reduce8(const float* part, int S, size_t split_stride, size_t off, float* acc) {
  float t[kLtMaxSplits][8];
#pragma unroll
  for (int j = 0; j < kLtMaxSplits; ++j)
    if (j < S) ldg256(part + (size_t)j * split_stride + off, t[j]);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = t[0][i];
#pragma unroll
  for (int j = 1; j < kLtMaxSplits; ++j)
    if (j < S) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __fadd_rn(acc[i], t[j][i]);
    }
}

// Tile shape and split count per GEMM (host): see layer_tail.cu.  max_bn caps the column tile.
bool lt_pick(int N, int K, int sms, bool split, LtGemm* out, int max_bn = 128);

constexpr int kLtCounterBytes = 2048;     // grid-barrier / exit counter region (zero between launches)

}  // namespace vlg
