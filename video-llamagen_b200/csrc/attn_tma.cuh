// Single-query attention over the bf16 KV-cache prefix [0, pos] -- the bandwidth part of the decode step
// (84 % of the algorithmic bytes at batch-64 CFG) -- written as a CTA-level PHASE: the stand-alone
// kernel (attn_tma.cu) runs it once per launch, the persistent decode-step kernel (decode_step.cu) runs it once per
// layer on the same shared-memory ring, carrying the ring / slot counters from phase to phase.
//
// HBM-bound streaming work: every K/V byte of the valid prefix is read exactly once per step and used for
// one FMA, so the design goal is bytes in flight, not flops:
//   * persistent CTAs (one per SM), work items = (row, head[, sequence split]) dealt round-robin;
//   * one TMA producer thread streams 32-position K and V tiles (4 KB each, 128-byte swizzle) through a
//     16-stage shared-memory ring (128 KB in flight per SM) and runs ahead ACROSS work items, so the pipe
//     never drains between the short (<= 19-tile) sequences of one step;
//   * 8 consumer warps each own every 8th tile: one lane per cache position for q.k (the swizzle makes the
//     16-byte row reads conflict-free), warp-level online softmax in the exp2 domain, then 8 lanes per
//     position for p.V; per-warp partial (m, l, o) states are merged by a dedicated merger warp once per item.
// Numerics: fp32 throughout, single rounding of the output to bf16 (torch math SDPA);
// the softmax scale is folded into q (1/sqrt(D) is a power of two for D = 64) together with log2(e).
#pragma once
#include "common.cuh"

namespace vlg {

constexpr int kTile = 32;                   // cache positions per TMA tile
constexpr int kCW = 8;                      // consumer warps
constexpr int kStagesA = 16;                // ring depth (multiple of kCW: stage s always belongs to warp s % kCW)
constexpr int kTileBytes = kTile * 128;     // one K or V tile (D = 64 bf16 = 128 B per position)
constexpr int kStageBytesA = 2 * kTileBytes;
constexpr int kAttnRingBytes = kStagesA * kStageBytesA;
constexpr int kMaxSplitsT = 32;

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

// Everything of the phase's shared state except the tile ring itself (which the persistent kernel shares with its GEMM
// phases).  SLOTS = in-flight work items per CTA whose partial states await their merge.
template <int SLOTS> struct AttnSide {
  ItemSlot slot[SLOTS][kCW];
  float p[kCW][kTile];
  float q[kCW][64];                         // QSMEM variant: the current item's scaled query, read back as warp broadcasts
  uint64_t full[kStagesA], empty[kStagesA];
  volatile unsigned cnt[SLOTS];             // warps that have published their partial for the item in this slot
  volatile unsigned gen[SLOTS];             // how many items have been merged out of this slot (slot reuse gate)
};

template <int SLOTS> __device__ __forceinline__ void attn_side_init(AttnSide<SLOTS>& sd) {   // one thread, before a CTA barrier
  for (int s = 0; s < kStagesA; ++s) { mbar_init(&sd.full[s], 1); mbar_init(&sd.empty[s], 1); }
  for (int s = 0; s < SLOTS; ++s) { sd.cnt[s] = 0; sd.gen[s] = 0; }
}

// What this CTA has pushed through the ring / the slots in earlier phases (identical in every thread).
struct AttnProgress { uint32_t tiles, items; };

enum AttnRole : int { ATTN_CONSUMER = 0, ATTN_PRODUCER = 1, ATTN_MERGER = 2, ATTN_IDLE = 3 };

struct AttnPhaseArgs {
  const __nv_bfloat16* q;     // [rows, H, 64]
  __nv_bfloat16* out;         // [rows, H * 64]
  float* part;                // split partials (n_splits > 1)
  int n_rh, H, S_max;
  int tile_row_base;          // first row of this layer's cache in the K/V tensor maps' row space
  int pos;
  const uint8_t* mask;        // [rows, T_prefix] or null
  int T_prefix;
  int n_splits;
};

// role: what this WARP does (whole warps only); cw: consumer index 0..kCW-1 for ATTN_CONSUMER.
// QSMEM: keep the scaled query in shared memory instead of 64 registers per consumer lane (the persistent kernel shares
// its register file with the GEMM phases' state and would otherwise spill inside the tile loop).
// TRACK: advance `prog` at the end (only the persistent kernel, which runs more than one phase, needs it).
template <int SLOTS, bool QSMEM, bool TRACK>
__device__ __forceinline__ void attn_tma_phase(uint8_t* ring, AttnSide<SLOTS>& sd, const TensorMap* tmK, const TensorMap* tmV,
                                               const AttnPhaseArgs& A, int role, int cw, AttnProgress& prog) {
  const int lane = threadIdx.x & 31;
  const int H = A.H, S_max = A.S_max, pos = A.pos, n_splits = A.n_splits, T_prefix = A.T_prefix;
  const uint8_t* mask = A.mask;
  const int n = pos + 1;
  int chunk = (n + n_splits - 1) / n_splits;
  chunk = (chunk + kTile - 1) / kTile * kTile;
  const int n_items = A.n_rh * n_splits;

  if (role == ATTN_PRODUCER) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      uint32_t gt = prog.tiles;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int rh = item / n_splits, split = item - rh * n_splits;
        const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
        const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
        const int row0 = A.tile_row_base + rh * S_max + s_begin;
        for (int t = 0; t < n_tiles; ++t, ++gt) {
          const uint32_t st = gt % kStagesA, it = gt / kStagesA;
          mbar_wait(&sd.empty[st], (it & 1) ^ 1);
          mbar_expect_tx(&sd.full[st], kStageBytesA);
          uint8_t* dst = ring + st * kStageBytesA;
          tma_load_2d_hint(dst, tmK, &sd.full[st], 0, row0 + t * kTile, kEvictFirst);
          tma_load_2d_hint(dst + kTileBytes, tmV, &sd.full[st], 0, row0 + t * kTile, kEvictFirst);
        }
      }
    }
    __syncwarp();
  } else if (role == ATTN_MERGER) {
    // ------------------------------------------------------------------ merger
    // Folds the consumers' partial (m, l, o) states of each multi-warp item and writes the output, so that no consumer
    // ever stalls its tile stream on merge work (otherwise the slowest warp of every item also pays for the merge).
    uint32_t gtm = prog.tiles;
    uint32_t mi = prog.items;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++mi) {
      const int rh = item / n_splits, split = item - rh * n_splits;
      const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
      const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
      const int w0 = (int)(gtm % kCW);
      const int participants = n_tiles < kCW ? n_tiles : kCW;
      gtm += n_tiles;
      if (participants <= 1) continue;                                // finalised by its only consumer
      const int sj = (int)(mi % SLOTS);
      const unsigned my_gen = mi / SLOTS;
      if (lane == 0) {
        uint32_t spins = 0;
        while (smem_ld_acquire(&sd.cnt[sj]) != (unsigned)participants)
          if (++spins > (1u << 26)) { printf("vlg: attention merge timeout\n"); __trap(); }
      }
      __syncwarp();
      const ItemSlot* sl = sd.slot[sj];
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
        *reinterpret_cast<uint32_t*>(A.out + (size_t)rh * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
      } else {
        float* pp = A.part + ((size_t)rh * n_splits + split) * 66;
        *reinterpret_cast<float2*>(pp + 2 + 2 * lane) = make_float2(o0, o1);
        if (lane == 0) { pp[0] = mm * 0.6931471805599453f; pp[1] = lt; }   // natural-log units for the combiner
      }
      __syncwarp();                                                   // all lanes are done with the slot
      if (lane == 0) {
        sd.cnt[sj] = 0;
        smem_st_release(&sd.gen[sj], my_gen + 1);                     // hand the slot to item mi + SLOTS
      }
    }
  } else if (role == ATTN_CONSUMER) {
    // ------------------------------------------------------------------ consumers
    const int warp = cw;
    const int g = lane >> 3, c = lane & 7;
    const uint32_t ring_base = smem_u32(ring), p_base = smem_u32(&sd.p[warp][0]);
    uint32_t gt0 = prog.tiles;                            // tiles issued before the current item
    uint32_t item_i = prog.items;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_i) {
      const int rh = item / n_splits, split = item - rh * n_splits;
      const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
      const int n_tiles = s_end > s_begin ? (s_end - s_begin + kTile - 1) / kTile : 0;
      const int w0 = (int)(gt0 % kCW);                                  // warp that owns the item's first tile
      const int participants = n_tiles < kCW ? n_tiles : kCW;           // warps (w0 + k) % kCW, k < participants
      const int first_t = (warp - w0 + kCW) % kCW;                      // first tile of this item owned by this warp
      const uint32_t gt_item = gt0;
      gt0 += n_tiles;
      if (!(n_tiles > 0 ? first_t < n_tiles : warp == 0)) continue;     // no share in this item: move on, no sync
      const uint8_t* mrow = mask ? mask + (size_t)(rh / H) * T_prefix : nullptr;

      float qf[QSMEM ? 1 : 64];
      const uint32_t q_base = smem_u32(&sd.q[warp][0]);
      {
        const float kScale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
        if constexpr (QSMEM) {
          const uint32_t qw = *reinterpret_cast<const uint32_t*>(A.q + (size_t)rh * 64 + 2 * lane);
          __syncwarp();                                       // previous item's readers of sd.q are done
          sts32f(q_base + lane * 8, bf16lo(qw) * kScale);
          sts32f(q_base + lane * 8 + 4, bf16hi(qw) * kScale);
          __syncwarp();
        } else {
          const uint4* qv = reinterpret_cast<const uint4*>(A.q + (size_t)rh * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 qr = qv[i];
            qf[(8 * i + 0) % (QSMEM ? 1 : 64)] = bf16lo(qr.x) * kScale; qf[(8 * i + 1) % (QSMEM ? 1 : 64)] = bf16hi(qr.x) * kScale;
            qf[(8 * i + 2) % (QSMEM ? 1 : 64)] = bf16lo(qr.y) * kScale; qf[(8 * i + 3) % (QSMEM ? 1 : 64)] = bf16hi(qr.y) * kScale;
            qf[(8 * i + 4) % (QSMEM ? 1 : 64)] = bf16lo(qr.z) * kScale; qf[(8 * i + 5) % (QSMEM ? 1 : 64)] = bf16hi(qr.z) * kScale;
            qf[(8 * i + 6) % (QSMEM ? 1 : 64)] = bf16lo(qr.w) * kScale; qf[(8 * i + 7) % (QSMEM ? 1 : 64)] = bf16hi(qr.w) * kScale;
          }
        }
      }

      float m = -INFINITY, l = 0.f, acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;

      for (int t = first_t; t < n_tiles; t += kCW) {
        const uint32_t gt = gt_item + t, st = gt % kStagesA, it = gt / kStagesA;
        mbar_wait(&sd.full[st], it & 1);
        const uint32_t kt = ring_base + st * kStageBytesA;
        const uint32_t vt = kt + kTileBytes;
        const int sidx = s_begin + t * kTile + lane;     // this lane's cache position

        // ---- scores: lane <-> position, 4 independent partial sums
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 kk = lds128(kt + lane * 128 + ((ch ^ (lane & 7)) << 4));
          float q8[8];
          if constexpr (QSMEM) {
            const float4 qa = lds128f(q_base + ch * 32), qb = lds128f(q_base + ch * 32 + 16);
            q8[0] = qa.x; q8[1] = qa.y; q8[2] = qa.z; q8[3] = qa.w; q8[4] = qb.x; q8[5] = qb.y; q8[6] = qb.z; q8[7] = qb.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) q8[i] = qf[(8 * ch + i) % (QSMEM ? 1 : 64)];
          }
          d0 = fmaf(q8[0], bf16lo(kk.x), d0); d1 = fmaf(q8[1], bf16hi(kk.x), d1);
          d2 = fmaf(q8[2], bf16lo(kk.y), d2); d3 = fmaf(q8[3], bf16hi(kk.y), d3);
          d0 = fmaf(q8[4], bf16lo(kk.z), d0); d1 = fmaf(q8[5], bf16hi(kk.z), d1);
          d2 = fmaf(q8[6], bf16lo(kk.w), d2); d3 = fmaf(q8[7], bf16hi(kk.w), d3);
        }
        float sc = (d0 + d1) + (d2 + d3);
        bool vis = sidx < s_end;
        if (mrow != nullptr && sidx < T_prefix) vis = vis && (sidx == pos || mrow[sidx] != 0);
        sc = vis ? sc : -INFINITY;

        // ---- online softmax (log2 domain)
        const float m_new = fmaxf(m, warp_max(sc));
        if (m_new == -INFINITY) {                         // whole tile masked and nothing seen yet
          __syncwarp();
          if (lane == 0) mbar_arrive(&sd.empty[st]);
          continue;
        }
        const float corr = exp2f(m - m_new);
        const float p = exp2f(sc - m_new);
        m = m_new;
        l = fmaf(l, corr, p);                             // per-lane partial of the row sum
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= corr;
        __syncwarp();                                     // previous tile's readers of sd.p are done
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
        if (lane == 0) mbar_arrive(&sd.empty[st]);
      }

      // ---- this warp's share of the item is done.  Warps flow through the items independently (no CTA barrier):
      // every participant publishes its partial (m, l, o); the merger warp folds them and writes the output.
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
          *reinterpret_cast<uint32_t*>(A.out + (size_t)rh * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
        } else {
          float* pp = A.part + ((size_t)rh * n_splits + split) * 66;
          *reinterpret_cast<float2*>(pp + 2 + 2 * lane) = make_float2(o0, o1);
          if (lane == 0) { pp[0] = m * 0.6931471805599453f; pp[1] = l; }
        }
        continue;
      }
      // multi-warp item: publish the partial state; the merger warp does the rest
      const int sj = (int)(item_i % SLOTS);
      const unsigned my_gen = item_i / SLOTS;
      if (lane == 0) {                                                  // slot reuse gate: previous occupant merged?
        uint32_t spins = 0;
        while (smem_ld_acquire(&sd.gen[sj]) != my_gen)
          if (++spins > (1u << 26)) { printf("vlg: attention slot timeout\n"); __trap(); }
      }
      __syncwarp();
      ItemSlot& pub = sd.slot[sj][warp];
      if (lane == 0) { pub.m = m; pub.l = l; }
      if (lane < 8) {
        *reinterpret_cast<float4*>(&pub.o[8 * c]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(&pub.o[8 * c + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      __syncwarp();                                                     // lanes' slot writes ordered before lane 0's release
      if (lane == 0) smem_atom_add_acq_rel(&sd.cnt[sj], 1u);
    }
  }

  // every thread advances the CTA's progress by what this phase pushed through the ring and the slots
  if constexpr (TRACK) {
    if ((int)blockIdx.x < n_items) {
      const uint32_t mine = (uint32_t)((n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1);
      if (n_splits == 1) {
        prog.tiles += mine * (uint32_t)((n + kTile - 1) / kTile);
      } else {
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
          const int split = item % n_splits;
          const int s_begin = split * chunk, s_end = min(n, s_begin + chunk);
          prog.tiles += s_end > s_begin ? (uint32_t)((s_end - s_begin + kTile - 1) / kTile) : 0u;
        }
      }
      prog.items += mine;
    }
  }
}

}  // namespace vlg
