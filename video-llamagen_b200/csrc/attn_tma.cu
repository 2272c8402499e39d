// Single-query attention over the bf16 KV-cache prefix [0, pos] -- the bandwidth kernel of the decode step
// (84 % of the algorithmic bytes at batch-64 CFG).
//
// HBM-bound streaming work: every K/V byte of the valid prefix is read exactly once per step and used for
// one FMA, so the design goal is bytes in flight, not flops:
//   * persistent CTAs (one per SM), work items = (row, head[, sequence split]) dealt round-robin;
//   * warp 0 is a TMA producer: it streams 32-position K and V tiles (4 KB each, 128-byte swizzle) through a
//     16-stage shared-memory ring (128 KB in flight per SM) and runs ahead ACROSS work items, so the pipe
//     never drains between the short (<= 19-tile) sequences of one step;
//   * 8 consumer warps each own every 8th tile: one lane per cache position for q.k (the swizzle makes the
//     16-byte row reads conflict-free), warp-level online softmax in the exp2 domain, then 8 lanes per
//     position for p.V; per-warp partial (m, l, o) states are merged through shared memory once per item.
// Numerics: fp32 throughout, single rounding of the output to bf16 (torch math SDPA);
// the softmax scale is folded into q (1/sqrt(D) is a power of two for D = 64) together with log2(e).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace vlg {

constexpr int kTile = 32;                   // cache positions per TMA tile
#ifndef VLG_ATTN_CW
#define VLG_ATTN_CW 8
#endif
constexpr int kCW = VLG_ATTN_CW;            // consumer warps
#ifndef VLG_ATTN_STAGES
#define VLG_ATTN_STAGES 16
#endif
#ifndef VLG_ATTN_SLOTS
#define VLG_ATTN_SLOTS 16
#endif
constexpr int kStagesA = VLG_ATTN_STAGES;    // ring depth (multiple of kCW: stage s always belongs to warp s % kCW)
constexpr int kTileBytes = kTile * 128;     // one K or V tile (D = 64 bf16 = 128 B per position)
constexpr int kStageBytesA = 2 * kTileBytes;
constexpr int kAttnTmaThreads = (kCW + 2) * 32;     // consumers, TMA producer (warp kCW), merger (warp kCW + 1)
constexpr int kMaxSplitsT = 32;

constexpr int kSlots = VLG_ATTN_SLOTS;       // in-flight work items per CTA whose partial states await their merge

// Explicit shared-space accesses: the ring lives behind a manually aligned pointer, which makes the compiler fall back
// to generic LD/ST (slower address-space resolution on every access) unless the state space is spelled out.
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v)); }

// CTA-scope synchronisation on shared-memory words (generic-address atomics would take the global atomic path)
__device__ __forceinline__ unsigned smem_atom_add_acq_rel(volatile unsigned* p, unsigned v) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(const_cast<unsigned*>(p))), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ unsigned smem_ld_acquire(const volatile unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(const_cast<const unsigned*>(p))) : "memory");
  return v;
}
__device__ __forceinline__ void smem_st_release(volatile unsigned* p, unsigned v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(const_cast<unsigned*>(p))), "r"(v) : "memory");
}

struct ItemSlot { float m, l, pad0, pad1; float o[64]; };

struct AttnTmaSmem {
  uint8_t ring[kStagesA][kStageBytesA];     // 1024-byte aligned (swizzle atom = 8 rows x 128 B)
  ItemSlot slot[kSlots][kCW];
  float p[kCW][kTile];
  float q[kCW][64];                         // the current item's scaled query of each consumer warp, read back as warp broadcasts
  uint64_t full[kStagesA], empty[kStagesA];
  volatile unsigned cnt[kSlots];            // warps that have published their partial for the item in this slot
  volatile unsigned gen[kSlots];            // how many items have been merged out of this slot (slot reuse gate)
};

__global__ void __launch_bounds__(kAttnTmaThreads, 1)
attn_decode_tma_kernel(const __grid_constant__ TensorMap tmK, const __grid_constant__ TensorMap tmV,
                       const __nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ out, float* __restrict__ part,
                       int n_rh, int H, int S_max, const int* pos_ptr, int pos_imm,
                       const uint8_t* __restrict__ mask, int T_prefix, int n_splits, int early) {
  extern __shared__ uint8_t smem_raw[];
  AttnTmaSmem& sm = *reinterpret_cast<AttnTmaSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int s = 0; s < kStagesA; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < kSlots; ++s) { sm.cnt[s] = 0; sm.gen[s] = 0; }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  // q and the K/V row of the CURRENT position come from the previous kernel; every older cache row was written in earlier
  // steps.  With `early` the producer starts streaming the first item's old tiles before it waits (the position word was
  // written by the previous step's sampler, several completed kernels ago -- see launch_attn_tma).
  uint32_t early_tiles = 0;
  if (warp == kCW) {
    if (early && elect_one()) {
      const int pos_e = pos_ptr ? *pos_ptr : pos_imm;
      const int n_e = pos_e + 1;
      int chunk_e = (n_e + n_splits - 1) / n_splits;
      chunk_e = (chunk_e + kTile - 1) / kTile * kTile;
      if ((int)blockIdx.x < n_rh * n_splits) {
        const int rh = blockIdx.x / n_splits, split = blockIdx.x - rh * n_splits;
        const int s_begin = split * chunk_e, s_end = min(n_e, s_begin + chunk_e);
        int n_old = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
        if (s_end == n_e) n_old -= 1;                    // the tile that holds the current position waits for the predecessor
        n_old = min(n_old, kStagesA);
        const int row0 = rh * S_max + s_begin;
        for (int t = 0; t < n_old; ++t) {
          mbar_expect_tx(&sm.full[t], kStageBytesA);
          tma_load_2d_hint(sm.ring[t], &tmK, &sm.full[t], 0, row0 + t * kTile, kEvictFirst);
          tma_load_2d_hint(sm.ring[t] + kTileBytes, &tmV, &sm.full[t], 0, row0 + t * kTile, kEvictFirst);
        }
        early_tiles = n_old > 0 ? (uint32_t)n_old : 0u;
      }
    }
    early_tiles = __reduce_max_sync(0xffffffffu, early_tiles);   // the elected lane's count to the whole warp
  }
  pdl_wait();

  const int pos = pos_ptr ? *pos_ptr : pos_imm;
  const int n = pos + 1;
  int chunk = (n + n_splits - 1) / n_splits;
  chunk = (chunk + kTile - 1) / kTile * kTile;
  const int n_items = n_rh * n_splits;

  if (warp == kCW) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t gt = early_tiles;                          // the first item's leading tiles are already in flight
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rh = item / n_splits, split = item - rh * n_splits;
        const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
        const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
        const int row0 = rh * S_max + s_begin;
        for (int t = item == (int)blockIdx.x ? (int)early_tiles : 0; t < n_tiles; ++t, ++gt) {
          const uint32_t st = gt % kStagesA, it = gt / kStagesA;
          mbar_wait(&sm.empty[st], (it & 1) ^ 1);
          mbar_expect_tx(&sm.full[st], kStageBytesA);
          tma_load_2d_hint(sm.ring[st], &tmK, &sm.full[st], 0, row0 + t * kTile, kEvictFirst);
          tma_load_2d_hint(sm.ring[st] + kTileBytes, &tmV, &sm.full[st], 0, row0 + t * kTile, kEvictFirst);
        }
      }
    }
    return;
  }

  if (warp == kCW + 1) {
    // ------------------------------------------------------------------ merger
    // Folds the consumers' partial (m, l, o) states of each multi-warp item and writes the output, so that no consumer
    // ever stalls its tile stream on merge work (otherwise the slowest warp of every item also pays for the merge).
    uint32_t gtm = 0;
    int mi = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++mi) {
      const int rh = item / n_splits, split = item - rh * n_splits;
      const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
      const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
      const int w0 = (int)(gtm % kCW);
      const int participants = n_tiles < kCW ? n_tiles : kCW;
      gtm += n_tiles;
      if (participants <= 1) continue;                                // finalised by its only consumer
      const int sj = mi % kSlots;
      const unsigned my_gen = (unsigned)(mi / kSlots);
      if (lane == 0) {
        uint32_t spins = 0;
        while (smem_ld_acquire(&sm.cnt[sj]) != (unsigned)participants)
          if (++spins > (1u << 26)) { printf("vlg: attention merge timeout\n"); __trap(); }
      }
      __syncwarp();
      const ItemSlot* sl = sm.slot[sj];
      float mm = -INFINITY;
      for (int k = 0; k < participants; ++k) mm = fmaxf(mm, sl[(w0 + k) % kCW].m);
      float lt = 0.f, o0 = 0.f, o1 = 0.f;
      for (int k = 0; k < participants; ++k) {
        const ItemSlot& ps = sl[(w0 + k) % kCW];
        const float a = ps.m == -INFINITY ? 0.f : exp2f(ps.m - mm);
        lt = fmaf(ps.l, a, lt);
        const float2 ov = *reinterpret_cast<const float2*>(&ps.o[2 * lane]);
        o0 = fmaf(ov.x, a, o0);
        o1 = fmaf(ov.y, a, o1);
      }
      if (n_splits == 1) {
        const float inv = 1.0f / lt;
        *reinterpret_cast<uint32_t*>(out + (size_t)rh * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
      } else {
        float* pp = part + ((size_t)rh * n_splits + split) * 66;
        *reinterpret_cast<float2*>(pp + 2 + 2 * lane) = make_float2(o0, o1);
        if (lane == 0) { pp[0] = mm * 0.6931471805599453f; pp[1] = lt; }   // natural-log units for the combiner
      }
      __syncwarp();                                                   // all lanes are done with the slot
      if (lane == 0) {
        sm.cnt[sj] = 0;
        smem_st_release(&sm.gen[sj], my_gen + 1);                     // hand the slot to item mi + kSlots
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int g = lane >> 3, c = lane & 7;
  const uint32_t ring_base = smem_u32(&sm.ring[0][0]), p_base = smem_u32(&sm.p[warp][0]);
  const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  uint32_t gt0 = 0;                                     // tiles issued before the current item
  int item_i = 0;
  // this lane's two query elements, fetched one item ahead: the L2 round trip of the query would otherwise sit in front
  // of every item's first tile (items are 1-19 tiles long, a warp owns 1-3 of them)
  uint32_t qw_next = 0;
  if ((int)blockIdx.x < n_items)
    qw_next = *reinterpret_cast<const uint32_t*>(q + (size_t)(blockIdx.x / n_splits) * 64 + 2 * lane);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
    const int rh = item / n_splits, split = item - rh * n_splits;
    const uint32_t qw = qw_next;
    if (item + (int)gridDim.x < n_items)
      qw_next = *reinterpret_cast<const uint32_t*>(q + (size_t)((item + gridDim.x) / n_splits) * 64 + 2 * lane);
    const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
    const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
    const int w0 = (int)(gt0 % kCW);                                  // warp that owns the item's first tile
    const int participants = n_tiles < kCW ? n_tiles : kCW;           // warps (w0 + k) % kCW, k < participants
    const int first_t = (warp - w0 + kCW) % kCW;                      // first tile of this item owned by this warp
    const uint32_t gt_item = gt0;
    gt0 += n_tiles;
    if (!(n_tiles > 0 ? first_t < n_tiles : warp == 0)) continue;     // no share in this item: move on, no sync
    const uint8_t* mrow = mask ? mask + (size_t)(rh / H) * T_prefix : nullptr;

    // The scaled query lives in shared memory, not in 64 registers per lane: 108 instead of 152 registers, and the
    // broadcast reads (one wavefront each) are cheaper than the register pressure they remove (measured 1-3 % faster).
    const uint32_t q_base = smem_u32(&sm.q[warp][0]);
    {
      __syncwarp();                                       // previous item's readers of sm.q are done
      sts32f(q_base + lane * 8, bf16lo(qw) * kScale);
      sts32f(q_base + lane * 8 + 4, bf16hi(qw) * kScale);
      __syncwarp();
    }

    float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;

    for (int t = first_t; t < n_tiles; t += kCW) {
      const uint32_t gt = gt_item + t, st = gt % kStagesA, it = gt / kStagesA;
      mbar_wait(&sm.full[st], it & 1);
      const uint32_t kt = ring_base + st * kStageBytesA;
      const uint32_t vt = kt + kTileBytes;
      const int sidx = s_begin + t * kTile + lane;     // this lane's cache position

      // ---- scores: lane <-> position, 4 independent partial sums
      float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 kk = lds128(kt + lane * 128 + ((ch ^ (lane & 7)) << 4));
        const float4 qa = lds128f(q_base + ch * 32), qb = lds128f(q_base + ch * 32 + 16);
        d0 = fmaf(qa.x, bf16lo(kk.x), d0); d1 = fmaf(qa.y, bf16hi(kk.x), d1);
        d2 = fmaf(qa.z, bf16lo(kk.y), d2); d3 = fmaf(qa.w, bf16hi(kk.y), d3);
        d0 = fmaf(qb.x, bf16lo(kk.z), d0); d1 = fmaf(qb.y, bf16hi(kk.z), d1);
        d2 = fmaf(qb.z, bf16lo(kk.w), d2); d3 = fmaf(qb.w, bf16hi(kk.w), d3);
      }
      float sc = (d0 + d1) + (d2 + d3);
      bool vis = sidx < s_end;
      if (mrow != nullptr && sidx < T_prefix) vis = vis && (sidx == pos || mrow[sidx] != 0);
      sc = vis ? sc : -INFINITY;

      // ---- online softmax (log2 domain)
      const float m_new = fmaxf(m, warp_max(sc));
      if (m_new == -INFINITY) {                         // whole tile masked and nothing seen yet
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[st]);
        continue;
      }
      const float corr = exp2f(m - m_new);
      const float p = exp2f(sc - m_new);
      m = m_new;
      l = fmaf(l, corr, p);                             // per-lane partial of the row sum
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= corr;
      __syncwarp();                                     // previous tile's readers of sm.p are done
      sts32f(p_base + lane * 4, p);
      __syncwarp();

      // ---- o += p . V : lane (g, c) covers rows 8g..8g+7, dims 8c..8c+7
      const float4 pa = lds128f(p_base + 32 * g);
      const float4 pb = lds128f(p_base + 32 * g + 16);
      const float pj[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      // Rows past the end of the prefix carry p = 0; their bytes are finite (the cache is zero-initialised once and only
      // ever holds finite values; TMA zero-fills out-of-bounds rows), so no branch is needed to keep them out.
      for (int j = 0; j < 8; ++j) {
        const int row = 8 * g + j;
        const uint4 vv = lds128(vt + row * 128 + ((c ^ (row & 7)) << 4));
        acc[0] = fmaf(pj[j], bf16lo(vv.x), acc[0]); acc[1] = fmaf(pj[j], bf16hi(vv.x), acc[1]);
        acc[2] = fmaf(pj[j], bf16lo(vv.y), acc[2]); acc[3] = fmaf(pj[j], bf16hi(vv.y), acc[3]);
        acc[4] = fmaf(pj[j], bf16lo(vv.z), acc[4]); acc[5] = fmaf(pj[j], bf16hi(vv.z), acc[5]);
        acc[6] = fmaf(pj[j], bf16lo(vv.w), acc[6]); acc[7] = fmaf(pj[j], bf16hi(vv.w), acc[7]);
      }
      __syncwarp();                                     // all lanes finished reading the stage
      if (lane == 0) mbar_arrive(&sm.empty[st]);
    }

    // ---- this warp's share of the item is done.  Warps flow through the items independently (no CTA barrier):
    // every participant publishes its partial (m, l, o); the last one to arrive merges and writes the output.
    l = warp_sum(l);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (participants <= 1) {
      // single-warp item (or an empty split): finalise here; lane L needs dims 2L, 2L+1 = acc[(2L) % 8 (+1)] of lane L / 4
      float o0 = 0.f, o1 = 0.f;
      const int src = lane >> 2, sel = (lane & 3) * 2;
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const float a0 = __shfl_sync(0xffffffffu, acc[i], src), a1 = __shfl_sync(0xffffffffu, acc[i + 1], src);
        if (sel == i) { o0 = a0; o1 = a1; }
      }
      if (n_splits == 1) {
        const float inv = 1.0f / l;
        *reinterpret_cast<uint32_t*>(out + (size_t)rh * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
      } else {
        float* pp = part + ((size_t)rh * n_splits + split) * 66;
        *reinterpret_cast<float2*>(pp + 2 + 2 * lane) = make_float2(o0, o1);
        if (lane == 0) { pp[0] = m * 0.6931471805599453f; pp[1] = l; }
      }
      continue;
    }
    // multi-warp item: publish the partial state; the merger warp does the rest
    const int sj = item_i % kSlots;
    const unsigned my_gen = (unsigned)(item_i / kSlots);
    if (lane == 0) {                                                  // slot reuse gate: previous occupant merged?
      uint32_t spins = 0;
      while (smem_ld_acquire(&sm.gen[sj]) != my_gen)
        if (++spins > (1u << 26)) { printf("vlg: attention slot timeout\n"); __trap(); }
    }
    __syncwarp();
    ItemSlot& pub = sm.slot[sj][warp];
    if (lane == 0) { pub.m = m; pub.l = l; }
    if (lane < 8) {
      *reinterpret_cast<float4*>(&pub.o[8 * c]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(&pub.o[8 * c + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncwarp();                                                     // lanes' slot writes ordered before lane 0's release
    if (lane == 0) smem_atom_add_acq_rel(&sm.cnt[sj], 1u);
  }
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

  // Sequence splits: a small cost model over the position mix of the run (mean prefix = half of max_pos, since a captured
  // graph fixes the split count for every step).  One CTA streams ~53 GB/s (measured: the kernel is bound by its
  // shared-memory / issue side at 7.9 TB/s over 148 SMs), an item costs ~0.3 us of fixed work, and any split count > 1
  // pays for the combine launch (~3.5 us with its kernel boundary) -- so e.g. 256 (row, head) pairs on 148 SMs stay
  // unsplit (two rounds of whole items beat four rounds of half items plus a combine), while 40 long pairs split 7-fold.
  const int sms = num_sms();
  int n_splits = 1;
  {
    const double item_us = 0.5 * (a.max_pos + 1) * 256.0 / 53e3, fixed_us = 0.3, combine_us = 3.5;
    double best = 1e30;
    const int max_s = (a.max_pos + 1 + kTile - 1) / kTile < kMaxSplitsT ? (a.max_pos + 1 + kTile - 1) / kTile : kMaxSplitsT;
    for (int sp = 1; sp <= (max_s < 1 ? 1 : max_s); ++sp) {
      const double rounds = (double)(((long)n_rh * sp + sms - 1) / sms);
      const double t = rounds * (item_us / sp + fixed_us) + (sp > 1 ? combine_us : 0.0);
      if (t < best - 1e-9) { best = t; n_splits = sp; }
    }
    static const int force = [] { const char* e = getenv("VLG_ATTN_SPLITS"); return e ? atoi(e) : 0; }();   // A/B
    if (force > 0) n_splits = force > max_s ? (max_s < 1 ? 1 : max_s) : force;
  }
  if (n_splits > 1)
    VLG_REQUIRE(a.scratch && a.scratch_bytes >= (size_t)n_rh * n_splits * 66 * sizeof(float),
                "attn(tma): scratch too small for %d splits", n_splits);
  const int smem = (int)sizeof(AttnTmaSmem) + 1024;
  static PerDeviceOnce attr_once;
  VLG_CUDA(attr_once([&] { return cudaFuncSetAttribute(attn_decode_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }));
  const int n_items = n_rh * n_splits;
  const int grid = n_items < sms ? n_items : sms;
  // Early streaming (before griddepcontrol.wait) reads the position word and cache rows that earlier decode steps wrote.
  // Only a captured graph orders those writes before this launch for certain (consecutive graph launches serialise);
  // on a plain stream the programmatic-launch chain lets this kernel's prologue run while a kernel two places back is
  // still executing, so the shortcut is off there.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool early = a.early_prefetch && cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusActive;
  VLG_CUDA(launch_kernel(attn_decode_tma_kernel, dim3(grid), dim3(kAttnTmaThreads), (size_t)smem, st, tmK, tmV,
                         (const __nv_bfloat16*)a.q, (__nv_bfloat16*)a.out, a.scratch, n_rh, a.H, a.S_max, a.pos_ptr,
                         a.pos_imm, a.prefix_mask, a.T_prefix, n_splits, early ? 1 : 0));
  VLG_LAUNCH_CHECK();
  if (n_splits > 1) {
    VLG_CUDA(launch_kernel(attn_tma_combine_kernel, dim3(n_rh), dim3(64), 0, st, (const float*)a.scratch,
                           (__nv_bfloat16*)a.out, n_splits));
    VLG_LAUNCH_CHECK();
  }
  return VLG_OK;
}

}  // namespace vlg
