// Stand-alone launch of the attention phase (attn_tma.cuh): one kernel per layer, used by the per-layer decode path, the
// vlg_attn_decode entry point and the roofline measurement of bench.py.
#include <stdlib.h>

#include "attn_tma.cuh"
#include "kernels.h"

namespace vlg {

constexpr int kAttnTmaThreads = (kCW + 2) * 32;     // consumers, TMA producer (warp kCW), merger (warp kCW + 1)
constexpr int kSlots = 16;

struct AttnTmaSmem {
  uint8_t ring[kAttnRingBytes];             // 1024-byte aligned (swizzle atom = 8 rows x 128 B)
  AttnSide<kSlots> side;
};

template <bool QSMEM>
__global__ void __launch_bounds__(kAttnTmaThreads, 1)
attn_decode_tma_kernel(const __grid_constant__ TensorMap tmK, const __grid_constant__ TensorMap tmV,
                       const __nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ out, float* __restrict__ part,
                       int n_rh, int H, int S_max, const int* pos_ptr, int pos_imm,
                       const uint8_t* __restrict__ mask, int T_prefix, int n_splits) {
  extern __shared__ uint8_t smem_raw[];
  AttnTmaSmem& sm = *reinterpret_cast<AttnTmaSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    attn_side_init(sm.side);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();                                // q, the newest K/V row and the position come from the previous kernels

  AttnPhaseArgs A;
  A.q = q; A.out = out; A.part = part;
  A.n_rh = n_rh; A.H = H; A.S_max = S_max;
  A.tile_row_base = 0;
  A.pos = pos_ptr ? *pos_ptr : pos_imm;
  A.mask = mask; A.T_prefix = T_prefix; A.n_splits = n_splits;
  AttnProgress prog{0u, 0u};
  const int role = warp < kCW ? ATTN_CONSUMER : (warp == kCW ? ATTN_PRODUCER : ATTN_MERGER);
  attn_tma_phase<kSlots, QSMEM, false>(sm.ring, sm.side, &tmK, &tmV, A, role, warp, prog);
}

// folds the split partials (same layout as the generic kernel's): out = sum_s o_s e^{m_s - M} / sum_s l_s e^{m_s - M}
__global__ void attn_tma_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, int n_splits) {
  const int rh = blockIdx.x, d = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const float* pp = part + (size_t)rh * n_splits * 66;
  float mm = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mm = fmaxf(mm, pp[s * 66]);
  float lt = 0.f, o = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float ms = pp[s * 66];
    const float a = (ms == -INFINITY) ? 0.f : expf(ms - mm);
    lt += pp[s * 66 + 1] * a;
    o += pp[s * 66 + 2 + d] * a;
  }
  out[(size_t)rh * 64 + d] = __float2bfloat16_rn(o / lt);
}

int launch_attn_tma(int dtype, const AttnArgs& a, cudaStream_t st) {
  VLG_REQUIRE(dtype == VLG_BF16, "attn(tma): bf16 only");
  VLG_REQUIRE(a.D == 64, "attn(tma): head_dim %d not supported (64)", a.D);
  VLG_REQUIRE(a.rows > 0 && a.H > 0 && a.S_max > 0, "attn(tma): bad shape");
  const int n_rh = a.rows * a.H;
  const uint64_t total_rows = (uint64_t)n_rh * a.S_max;
  VLG_REQUIRE(total_rows < (1ull << 31), "attn(tma): cache too large for 32-bit TMA coordinates");
  TensorMap tmK, tmV;
  int rc = make_tmap_bf16_2d(&tmK, a.k_cache, total_rows, 64, 128, kTile, 64, true);
  if (rc) return rc;
  if ((rc = make_tmap_bf16_2d(&tmV, a.v_cache, total_rows, 64, 128, kTile, 64, true))) return rc;

  // sequence splits only when (row, head) pairs alone cannot occupy the SMs
  const int sms = num_sms();
  int n_splits = 1;
  if (n_rh < 2 * sms) {
    int want = (2 * sms + n_rh - 1) / n_rh, by_len = (a.max_pos + 1 + 2 * kTile - 1) / (2 * kTile);
    n_splits = want < by_len ? want : by_len;
    if (n_splits < 1) n_splits = 1;
    if (n_splits > kMaxSplitsT) n_splits = kMaxSplitsT;
  }
  if (n_splits > 1)
    VLG_REQUIRE(a.scratch && a.scratch_bytes >= (size_t)n_rh * n_splits * 66 * sizeof(float),
                "attn(tma): scratch too small for %d splits", n_splits);
  const int smem = (int)sizeof(AttnTmaSmem) + 1024;
  // query staged in shared memory (108 registers, measured 1-3 % faster than 152 registers with q in registers);
  // VLG_ATTN_QSMEM=0 selects the register variant for A/B runs
  static const bool qsmem = [] { const char* e = getenv("VLG_ATTN_QSMEM"); return !(e && e[0] == '0'); }();
  auto kern = qsmem ? attn_decode_tma_kernel<true> : attn_decode_tma_kernel<false>;
  static bool attr_set = false;
  if (!attr_set) {
    VLG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int n_items = n_rh * n_splits;
  const int grid = n_items < sms ? n_items : sms;
  VLG_CUDA(launch_kernel(kern, dim3(grid), dim3(kAttnTmaThreads), (size_t)smem, st, tmK, tmV,
                         (const __nv_bfloat16*)a.q, (__nv_bfloat16*)a.out, a.scratch, n_rh, a.H, a.S_max, a.pos_ptr,
                         a.pos_imm, a.prefix_mask, a.T_prefix, n_splits));
  VLG_LAUNCH_CHECK();
  if (n_splits > 1) {
    VLG_CUDA(launch_kernel(attn_tma_combine_kernel, dim3(n_rh), dim3(64), 0, st, (const float*)a.scratch,
                           (__nv_bfloat16*)a.out, n_splits));
    VLG_LAUNCH_CHECK();
  }
  return VLG_OK;
}

}  // namespace vlg
