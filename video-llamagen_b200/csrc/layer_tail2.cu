// Layer tail, second form: the same arithmetic as layer_tail.cu
//
//   wo(attn) + residual -> RMSNorm -> w1|w3 -> SwiGLU -> w2 + residual -> RMSNorm -> wqkv(next layer) -> RoPE + KV append
//   (gpt.py:255-259 TransformerBlock, :143-148 RMSNorm, :151-167 FeedForward, :207-227 Attention up to the cache write)
//
// but without CTA-wide grid barriers.  layer_tail.cu spends half of its 27 us per launch in six dependent exchanges through
// L2 (store -> release -> counter -> poll -> first TMA round trip: ~2 us each) during which every SM idles, and
// tools/micro/ingest_bench.cu shows that the rest is bounded by what one SM can pull out of L2 (~110 GB/s per SM whether the
// bytes are shared, private or multicast), so neither multicast nor more CTAs help.  What helps is keeping the SM's port busy
// while an exchange is in flight.  Here the rows are split into two GROUPS (rows [0, 64) and [64, 128) at cfg-2) that walk
// the six stages one after the other, and the CTA is a set of actors that only meet through mbarriers:
//
//   P   (warp 0, one thread)  TMA producer.  Weight tiles go into a W ring as soon as a slot is free (a whole stage ahead);
//                             a group's activation tiles go into the A ring once the counter of the (stage, group) that
//                             produces them shows every CTA's contribution.
//   M   (warp 1, one thread)  tcgen05.mma issuer: units in order (stage, group).  Group g accumulates in TMEM columns
//                             [128 g, 128 g + BN); its 64 rows sit on TMEM lanes [64 g, 64 g + 64).  A weight slot is
//                             released after the LAST group has used it: weights cross the SM's port once, like before.
//   E_g (4 warps per group)   the two warps pairs that may touch lanes [64 g, 64 g + 64): epilogues of group g's units
//                             (split-K partial store / SwiGLU / RoPE + cache append) and group g's row phases (fixed-order
//                             split-K reduction + residual + RMSNorm, one row per CTA), each followed by ONE release-add on
//                             the counter of that (stage, group).
//
// While group 0's exchange is in flight the port, the tensor pipe and the other group's warps work on group 1, and vice
// versa; the only waits left on the critical path are the dependent exchanges of ONE half-sized group.  The MMA always
// spans 128 lanes: a group's 64-row activation tile is placed so that its rows land on the group's lanes, the other 64
// lanes multiply whatever the neighbouring 8 KB of shared memory holds and are never read.
//
// Rounding points, split-K order and the row-phase arithmetic are those of layer_tail.cu (same helpers): for dim <= 1024
// the two kernels are bit-identical.
#include <algorithm>
#include "lt_common.cuh"

namespace vlg {

constexpr int kT2Threads = 320;
constexpr int kT2Slot = 8192;         // one ring slot: 64 rows x 64 k bf16 (activation tile of a group, or a weight tile of <= 64 rows)
constexpr int kT2NW = 20;             // W ring slots (slots 0..NW-2 in front of the A ring, slot NW-1 behind it, so that the 8 KB
constexpr int kT2NA = 8;              //   next to any A stage -- the lanes nobody reads -- are always inside the allocation)
constexpr int kT2TmemCols = 256;
constexpr int kT2CntStride = 32;      // words between counters: one 128-byte line each
enum { T2_WO = 0, T2_N1 = 1, T2_W13 = 2, T2_W2 = 3, T2_N2 = 4, T2_QKV = 5, T2_EXIT = 12 };
constexpr int t2_smem_bytes() { return (kT2NW + kT2NA) * kT2Slot + 1024 /*align slack*/ + 1024 /*barriers*/; }

struct T2Maps { TensorMap attn, xn, ffn, wo, w13, w2, wqkv; };

struct T2Params {
  int rows, dim, hidden;
  int has_front, has_qkv;
  int G;                     // row groups (1 or 2)
  int row0[2], nr[2];        // first row and row count of each group
  uint32_t a_bytes;          // bytes of one activation box (round8(nr[0]) x 128)
  LtGemm g_wo, g_w13, g_w2, g_qkv;
  float* part;               // split-K partials [S][rows][N] fp32
  __nv_bfloat16* h;
  __nv_bfloat16* xn;
  const __nv_bfloat16* norm1_w;
  const __nv_bfloat16* norm2_w;
  float eps;
  EpiParams ffn, rope;
  unsigned* cnt;             // counters [(stage * 2 + group) * kT2CntStride], exit counter at T2_EXIT; zero between launches
  volatile unsigned* err;
  int dbg, fast_silu;
};

struct T2Smem {
  uint64_t w_full[kT2NW], w_empty[kT2NW], a_full[kT2NA], a_empty[kT2NA];
  uint64_t acc_ready[2], acc_free[2];
  uint32_t tmem_slot;
  float red[2][4];
};

// globaltimer stamps, VLG_LT_DEBUG=1 (tools/lt_phases.py): [cta][slot]; slot 0 = start, 1 = past pdl_wait,
// 2 + stage * 2 + group = E_g finished that stage, 16 + stage * 2 + group = P saw that stage's input counter complete;
// 32 + 8 * group + i: the w13 stage in detail (M: 0 accumulator free, 1 first tile pair landed, 2 last MMA issued; E: 3 accumulator
// ready, 4 epilogue stored, 5 signalled; P: 6 first, 7 last activation tile requested)
__device__ unsigned long long g_lt2_ts[160 * 64];
__device__ __forceinline__ void t2_stamp(const T2Params& P, int slot) {
  if (P.dbg && P.has_front && P.has_qkv && blockIdx.x < 160) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_lt2_ts[blockIdx.x * 64 + slot] = t;
  }
}

__device__ __forceinline__ uint8_t* t2_w_slot(uint8_t* ring, uint32_t s) { return ring + (size_t)(s < kT2NW - 1 ? s : s + kT2NA) * kT2Slot; }
__device__ __forceinline__ uint8_t* t2_a_slot(uint8_t* ring, uint32_t s) { return ring + (size_t)(kT2NW - 1 + s) * kT2Slot; }
__device__ __forceinline__ unsigned* t2_cnt(const T2Params& P, int stage, int g) { return P.cnt + (stage * 2 + g) * kT2CntStride; }
__device__ __forceinline__ void t2_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// Wait until every CTA has added its contribution to a counter.  Bounded like the grid barrier of layer_tail.cu: if the grid
// is not co-resident the wait gives up after 2 s, sets the host-visible error word, and every later wait falls through.
__device__ __forceinline__ void t2_poll(const unsigned* ctr, volatile unsigned* err) {
  const unsigned target = gridDim.x;
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
__device__ __forceinline__ void t2_signal(unsigned* ctr) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory"); }

// ---------------------------------------------------------------------------------------------------- P: TMA producer
struct T2Prod {
  uint32_t gw, ga;
  int w_ahead;    // weight tiles of the unit whose activation tiles come next that are already requested
};

__device__ __forceinline__ void t2_w_tile(const LtGemm& G, const LtUnit& u, const TensorMap* tmB, uint8_t* ring, T2Smem& sm, uint32_t i, int kb) {
  const uint32_t s = i % kT2NW, it = i / kT2NW;
  mbar_wait(&sm.w_empty[s], (it & 1) ^ 1);
  mbar_expect_tx(&sm.w_full[s], G.BN * 128);
  tma_load_2d_hint(t2_w_slot(ring, s), tmB, &sm.w_full[s], (u.kb0 + kb) * 64, u.n0, kEvictFirst);
}
// Weight tiles of a unit, ahead of its activation tiles: as many as the ring can hold without this unit's own progress (the
// slots they wait for belong to earlier units, whose activation tiles are already on their way).  A unit deeper than the
// ring (K > 64 * kT2NW, one group only) gets the rest just in time from t2_issue_a.
__device__ __forceinline__ void t2_issue_w(const LtGemm& G, const TensorMap* tmB, uint8_t* ring, T2Smem& sm, T2Prod& p) {
  const LtUnit u = lt_unit(G, blockIdx.x);
  if (!u.valid) return;
  const int n = u.nkb < kT2NW ? u.nkb : kT2NW;
  for (int kb = 0; kb < n; ++kb) t2_w_tile(G, u, tmB, ring, sm, p.gw + kb, kb);
  p.w_ahead = n;
}
__device__ __forceinline__ void t2_issue_a(const LtGemm& G, const TensorMap* tmA, const TensorMap* tmB, int row0, uint32_t a_bytes, uint8_t* ring,
                                           T2Smem& sm, T2Prod& p) {
  const LtUnit u = lt_unit(G, blockIdx.x);
  if (!u.valid) return;
  for (int kb = 0; kb < u.nkb; ++kb) {
    if (kb >= p.w_ahead) { t2_w_tile(G, u, tmB, ring, sm, p.gw + kb, kb); p.w_ahead = kb + 1; }
    const uint32_t i = p.ga + kb, s = i % kT2NA, it = i / kT2NA;
    mbar_wait(&sm.a_empty[s], (it & 1) ^ 1);
    mbar_expect_tx(&sm.a_full[s], a_bytes);
    tma_load_2d_hint(t2_a_slot(ring, s), tmA, &sm.a_full[s], (u.kb0 + kb) * 64, row0, kEvictLast);
  }
  p.ga += u.nkb;
}
// activation tiles of one GEMM stage, group by group, each once its input is complete (dep < 0: input of the launch)
__device__ __forceinline__ void t2_stage_a(const T2Params& P, const LtGemm& G, const TensorMap* tmA, const TensorMap* tmB, int dep, uint8_t* ring,
                                           T2Smem& sm, T2Prod& p) {
  const LtUnit u = lt_unit(G, blockIdx.x);
  if (!u.valid) return;                                 // nothing to load: the counters are somebody else's business
  for (int g = 0; g < P.G; ++g) {
    if (dep >= 0) {
      t2_poll(t2_cnt(P, dep, g), P.err);
      t2_stamp(P, 16 + dep * 2 + g);
    }
    asm volatile("fence.proxy.async;" ::: "memory");    // other SMs' generic-proxy stores of A -> this SM's TMA reads
    if (P.dbg && dep == T2_N1) t2_stamp(P, 32 + 8 * g + 6);
    t2_issue_a(G, tmA, tmB, P.row0[g], P.a_bytes, ring, sm, p);
    if (P.dbg && dep == T2_N1) t2_stamp(P, 32 + 8 * g + 7);
  }
  p.gw += u.nkb;
  p.w_ahead = 0;
}

// ---------------------------------------------------------------------------------------------------- M: MMA issuer
struct T2Mma { uint32_t gw, ga, uses[2]; };

__device__ __forceinline__ void t2_mma_stage(const T2Params& P, const LtGemm& G, uint8_t* ring, T2Smem& sm, uint32_t tmem_base, T2Mma& m) {
  const LtUnit u = lt_unit(G, blockIdx.x);
  if (!u.valid) return;
  const uint32_t idesc = umma_idesc_bf16(128, G.BN);
  const bool detail = P.dbg && &G == &P.g_w13;
  for (int g = 0; g < P.G; ++g) {
    mbar_wait(&sm.acc_free[g], (m.uses[g] & 1) ^ 1);   // E_g has drained the previous unit's accumulator
    tc_fence_after();
    if (detail) t2_stamp(P, 32 + 8 * g + 0);
    const uint32_t acc = tmem_base + 128u * g;
    for (int kb = 0; kb < u.nkb; ++kb) {
      const uint32_t iw = m.gw + kb, sw = iw % kT2NW, ia = m.ga + kb, sa = ia % kT2NA;
      mbar_wait(&sm.w_full[sw], (iw / kT2NW) & 1);     // second group: the phase completed long ago, passes at once
      mbar_wait(&sm.a_full[sa], (ia / kT2NA) & 1);
      tc_fence_after();
      if (detail && kb == 0) t2_stamp(P, 32 + 8 * g + 1);
      // the group's 64 rows must land on lanes [64 g, 64 g + 64): start the 128-row operand 8 KB early for group 1
      const uint64_t adesc = umma_desc_sw128(smem_u32(t2_a_slot(ring, sa)) - (uint32_t)g * kT2Slot);
      const uint64_t bdesc = umma_desc_sw128(smem_u32(t2_w_slot(ring, sw)));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
      umma_commit(&sm.a_empty[sa]);
      if (g == P.G - 1) umma_commit(&sm.w_empty[sw]);
    }
    umma_commit(&sm.acc_ready[g]);
    if (detail) t2_stamp(P, 32 + 8 * g + 2);
    m.ga += u.nkb;
    m.uses[g] += 1;
  }
  m.gw += u.nkb;
}

// ---------------------------------------------------------------------------------------------------- E_g: epilogues, row phases
struct T2Epi {
  int g;          // group
  int wq;         // warp within the group: (quadrant & 1) + 2 * column half
  int lane;
  uint32_t uses;  // accumulator uses of this group so far
};

// MODE: -1 = store the fp32 partial tile (split-K); EPI_SWIGLU / EPI_QKV_ROPE = elementwise tail straight from TMEM
template <int MODE>
__device__ __forceinline__ void t2_gemm_epi(const T2Params& P, const LtGemm& G, const EpiParams& E, int stage, T2Smem& sm, uint32_t tmem_base, T2Epi& e) {
  const LtUnit u = lt_unit(G, blockIdx.x);
  if (u.valid) {
    const int quad = 2 * e.g + (e.wq & 1), half = e.wq >> 1;
    const int rg = (e.wq & 1) * 32 + e.lane;             // row within the group
    const int r = P.row0[e.g] + rg;
    const bool live = rg < P.nr[e.g];
    const uint32_t lane_base = tmem_base + 128u * e.g + (static_cast<uint32_t>(quad * 32) << 16);
    mbar_wait(&sm.acc_ready[e.g], e.uses & 1);
    tc_fence_after();
    if (P.dbg && stage == T2_W13 && e.wq == 0 && e.lane == 0) t2_stamp(P, 32 + 8 * e.g + 3);
    if constexpr (MODE < 0) {
      float* dst = P.part + ((size_t)u.split * P.rows + r) * G.N + u.n0;
      for (int c = half * 16; c < G.BN; c += 32) {
        float v[16];
        tmem_ld16(lane_base + c, v);
        if (live) {
          stg256(dst + c, v);
          stg256(dst + c + 8, v + 8);
        }
      }
    } else if constexpr (MODE == EPI_SWIGLU) {          // gpt.py:167 -- w1/w3 rows interleaved: columns (2j, 2j+1) = (w1 x, w3 x)_j
      for (int c = half * 32; c < G.BN; c += 64) {
        float v[32];
        tmem_ld16(lane_base + c, v);
        tmem_ld16(lane_base + c + 16, v + 16);
        if (live) {
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float a = rnd<__nv_bfloat16>(v[2 * i]), b = rnd<__nv_bfloat16>(v[2 * i + 1]);
            const float sl = P.fast_silu ? __fdividef(a, 1.0f + __expf(-a)) : silu_f(a);
            o[i] = __fmul_rn(rnd<__nv_bfloat16>(sl), b);
          }
          stg256_bf16x16(static_cast<__nv_bfloat16*>(E.out) + (size_t)r * E.ld_out + ((u.n0 + c) >> 1), o);
        }
      }
    } else {                                             // EPI_QKV_ROPE on 16 columns (a quarter head): gpt.py:215-227, 423-433, 177-185
      const int pos = E.pos_ptr ? *E.pos_ptr : E.pos_imm;
      for (int c = half * 16; c < G.BN; c += 32) {
        float v[16];
        tmem_ld16(lane_base + c, v);
        if (live) {
          const int n = u.n0 + c, sec = n / E.dim, cc = n - sec * E.dim, head = cc / E.D, dd = cc - head * E.D;
          float o[16];
          if (sec < 2) {
            const float2* tab = reinterpret_cast<const float2*>(E.rope) + (size_t)pos * (E.D >> 1) + (dd >> 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 cs = tab[i];
              const float x0 = rnd<__nv_bfloat16>(v[2 * i]), x1 = rnd<__nv_bfloat16>(v[2 * i + 1]);
              o[2 * i] = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
              o[2 * i + 1] = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = v[i];
          }
          __nv_bfloat16* dst;
          if (sec == 0) dst = static_cast<__nv_bfloat16*>(E.q_out) + (size_t)r * E.dim + cc;
          else dst = static_cast<__nv_bfloat16*>(sec == 1 ? E.k_cache : E.v_cache) + (((size_t)r * E.H + head) * E.S_max + pos) * E.D + dd;
          stg256_bf16x16(dst, o);
        }
      }
    }
    tc_fence_before();
    e.uses += 1;
  }
  t2_bar(e.g);                                           // the group's stores and TMEM reads are done
  if (e.wq == 0 && e.lane == 0) {
    if (u.valid) mbar_arrive(&sm.acc_free[e.g]);
    if (P.dbg && stage == T2_W13 && u.valid) t2_stamp(P, 32 + 8 * e.g + 4);
    t2_signal(t2_cnt(P, stage, e.g));
    if (P.dbg && stage == T2_W13 && u.valid) { __threadfence(); t2_stamp(P, 32 + 8 * e.g + 5); }
    t2_stamp(P, 2 + stage * 2 + e.g);
  }
}

// h <- T(h + T(sum of partials))  (when Gs != null), then xn <- T(T(h * rsqrt(mean(h^2) + eps)) * w): the group's rows, one
// per CTA; group 1's rows go to the CTAs at the far end of the grid so that no CTA owns a row of both groups.
__device__ __forceinline__ void t2_norm(const T2Params& P, const LtGemm* Gs, const __nv_bfloat16* w, int stage, int dep, T2Smem& sm, T2Epi& e) {
  const int t = e.wq * 32 + e.lane;                      // 0..127 within the group
  const int nvec = P.dim / 8;
  const int n_rows = P.nr[e.g];
  const int first = e.g == 0 ? (int)blockIdx.x : (int)((blockIdx.x + n_rows) % gridDim.x);
  if (first < n_rows) {
    if (dep >= 0) {
      if (t == 0) t2_poll(t2_cnt(P, dep, e.g), P.err);
      t2_bar(e.g);
    }
    for (int i = first; i < n_rows; i += gridDim.x) {
      const int r = P.row0[e.g] + i;
      float hn[2][8];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int v = t + j * 128;
        if (v < nvec) {
          __nv_bfloat16* hp = P.h + (size_t)r * P.dim + v * 8;
          const uint4 hu = ldg128_cg(hp);
          hn[j][0] = bf16lo(hu.x); hn[j][1] = bf16hi(hu.x); hn[j][2] = bf16lo(hu.y); hn[j][3] = bf16hi(hu.y);
          hn[j][4] = bf16lo(hu.z); hn[j][5] = bf16hi(hu.z); hn[j][6] = bf16lo(hu.w); hn[j][7] = bf16hi(hu.w);
          if (Gs != nullptr) {
            float acc[8];
            lt_reduce8(P.part, Gs->S, (size_t)P.rows * Gs->N, (size_t)r * Gs->N + v * 8, acc);
#pragma unroll
            for (int k = 0; k < 8; ++k) hn[j][k] = rnd<__nv_bfloat16>(__fadd_rn(hn[j][k], rnd<__nv_bfloat16>(acc[k])));
            store8<__nv_bfloat16>(hp, hn[j]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) ss = __fadd_rn(ss, __fmul_rn(hn[j][k], hn[j][k]));
        }
      }
      ss = warp_sum(ss);
      if (e.lane == 0) sm.red[e.g][e.wq] = ss;
      t2_bar(e.g);
      const float tot = ((sm.red[e.g][0] + sm.red[e.g][1]) + sm.red[e.g][2]) + sm.red[e.g][3];
      const float rs = rsqrtf(__fdiv_rn(tot, (float)P.dim) + P.eps);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int v = t + j * 128;
        if (v < nvec) {
          float g8[8], o[8];
          load8<__nv_bfloat16>(w + v * 8, g8);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = __fmul_rn(rnd<__nv_bfloat16>(__fmul_rn(hn[j][k], rs)), g8[k]);
          store8<__nv_bfloat16>(P.xn + (size_t)r * P.dim + v * 8, o);
        }
      }
      if (i + (int)gridDim.x < n_rows) t2_bar(e.g);      // red[] is reused by the next row
    }
  }
  t2_bar(e.g);
  if (t == 0) {
    t2_signal(t2_cnt(P, stage, e.g));
    t2_stamp(P, 2 + stage * 2 + e.g);
  }
}

// ---------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kT2Threads, 1)
layer_tail2_kernel(const __grid_constant__ T2Maps tm, const T2Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  T2Smem& sm = *reinterpret_cast<T2Smem*>(ring + (kT2NW + kT2NA) * kT2Slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (P.has_front) { tma_prefetch_desc(&tm.attn); tma_prefetch_desc(&tm.ffn); tma_prefetch_desc(&tm.wo); tma_prefetch_desc(&tm.w13); tma_prefetch_desc(&tm.w2); }
    tma_prefetch_desc(&tm.xn);
    if (P.has_qkv) tma_prefetch_desc(&tm.wqkv);
    for (int s = 0; s < kT2NW; ++s) { mbar_init(&sm.w_full[s], 1); mbar_init(&sm.w_empty[s], 1); }
    for (int s = 0; s < kT2NA; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], 1); }
    for (int g = 0; g < 2; ++g) { mbar_init(&sm.acc_ready[g], 1); mbar_init(&sm.acc_free[g], 1); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kT2TmemCols>(&sm.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_slot;
  if (threadIdx.x == 0) t2_stamp(P, 0);
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ P
    if (elect_one()) {
      T2Prod p{0, 0, 0};
      // the first GEMM's weight tiles do not depend on the previous kernel: request them before waiting for it
      if (P.has_front) t2_issue_w(P.g_wo, &tm.wo, ring, sm, p);
      else if (P.has_qkv) t2_issue_w(P.g_qkv, &tm.wqkv, ring, sm, p);
      pdl_wait();
      t2_stamp(P, 1);
      if (P.has_front) {
        t2_stage_a(P, P.g_wo, &tm.attn, &tm.wo, -1, ring, sm, p);
        t2_issue_w(P.g_w13, &tm.w13, ring, sm, p);
        t2_stage_a(P, P.g_w13, &tm.xn, &tm.w13, T2_N1, ring, sm, p);
        t2_issue_w(P.g_w2, &tm.w2, ring, sm, p);
        t2_stage_a(P, P.g_w2, &tm.ffn, &tm.w2, T2_W13, ring, sm, p);
        if (P.has_qkv) t2_issue_w(P.g_qkv, &tm.wqkv, ring, sm, p);
      }
      if (P.has_qkv) t2_stage_a(P, P.g_qkv, &tm.xn, &tm.wqkv, T2_N2, ring, sm, p);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ M
    if (elect_one()) {
      T2Mma m{0, 0, {0, 0}};
      if (P.has_front) {
        t2_mma_stage(P, P.g_wo, ring, sm, tmem_base, m);
        t2_mma_stage(P, P.g_w13, ring, sm, tmem_base, m);
        t2_mma_stage(P, P.g_w2, ring, sm, tmem_base, m);
      }
      if (P.has_qkv) t2_mma_stage(P, P.g_qkv, ring, sm, tmem_base, m);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ E_0 (TMEM quadrants 0, 1) and E_1 (quadrants 2, 3)
    const int quad = warp & 3;
    T2Epi e{quad >> 1, (quad & 1) + 2 * ((warp - 2) >> 2), lane, 0};
    if (e.g < P.G) {
      pdl_wait();
      if (P.has_front) {
        t2_gemm_epi<-1>(P, P.g_wo, P.ffn, T2_WO, sm, tmem_base, e);
        t2_norm(P, &P.g_wo, P.norm1_w, T2_N1, T2_WO, sm, e);
        t2_gemm_epi<EPI_SWIGLU>(P, P.g_w13, P.ffn, T2_W13, sm, tmem_base, e);
        t2_gemm_epi<-1>(P, P.g_w2, P.ffn, T2_W2, sm, tmem_base, e);
        t2_norm(P, &P.g_w2, P.norm2_w, T2_N2, T2_W2, sm, e);
      } else {
        t2_norm(P, nullptr, P.norm2_w, T2_N2, -1, sm, e);
      }
      if (P.has_qkv) t2_gemm_epi<EPI_QKV_ROPE>(P, P.g_qkv, P.rope, T2_QKV, sm, tmem_base, e);
    }
  }

  // leave the counters at zero for the next launch: the last CTA out has seen everyone finish every wait
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(P.cnt + T2_EXIT * kT2CntStride, 1u);
    if (old == gridDim.x - 1) {
      for (int i = 0; i <= T2_EXIT; ++i) P.cnt[i * kT2CntStride] = 0;
      __threadfence();
    }
  }
  if (warp == 1) tmem_dealloc<kT2TmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------- host
namespace {
int t2_groups_env() { static const int v = [] { const char* e = getenv("VLG_LT_GROUPS"); return e ? atoi(e) : 0; }(); return v; }

// tile shapes (column tiles <= 64 wide: a weight tile is one ring slot) and the group count for this row count
bool t2_plan(int rows, int dim, int hidden, LtGemm* wo, LtGemm* w13, LtGemm* w2, LtGemm* qkv, int* groups) {
  if (rows <= 0 || rows > 128 || dim % 64 != 0 || hidden % 64 != 0 || dim > 2048) return false;
  const int sms = num_sms();
  if (!lt_pick(dim, dim, sms, true, wo, 64) || !lt_pick(2 * hidden, dim, sms, false, w13, 64) || !lt_pick(dim, hidden, sms, true, w2, 64) ||
      !lt_pick(3 * dim, dim, sms, false, qkv, 64))
    return false;
  // two groups keep a unit's weight tiles resident until the second group has used them: the whole unit must fit the W ring
  const int max_kb = std::max(std::max(wo->kb_per, w13->kb_per), std::max(w2->kb_per, qkv->kb_per));
  const bool fits2 = max_kb <= kT2NW;
  int G = rows > 64 ? 2 : 1;
  const int want = t2_groups_env();
  if (want == 2 && rows >= 16) G = 2;
  if (want == 1 && rows <= 64) G = 1;
  if (G == 2 && !fits2) {
    if (rows > 64) return false;
    G = 1;
  }
  *groups = G;
  return true;
}
}  // namespace

bool layer_tail2_supported(int rows, int dim, int hidden) {
  static const int ver = [] { const char* e = getenv("VLG_LT_V"); return e ? atoi(e) : 1; }();   // measured slower (see the header): off unless asked for
  if (ver != 2) return false;
  LtGemm a, b, c, d;
  int G;
  return t2_plan(rows, dim, hidden, &a, &b, &c, &d, &G);
}

int launch_layer_tail2(const LtLaunch& a, cudaStream_t st) {
  T2Params P{};
  VLG_REQUIRE(t2_plan(a.rows, a.dim, a.hidden, &P.g_wo, &P.g_w13, &P.g_w2, &P.g_qkv, &P.G), "layer_tail2: unsupported shape rows=%d dim=%d hidden=%d",
              a.rows, a.dim, a.hidden);
  VLG_REQUIRE(a.wo || a.wqkv, "layer_tail2: nothing to do");
  P.rows = a.rows; P.dim = a.dim; P.hidden = a.hidden;
  P.has_front = a.wo != nullptr;
  P.has_qkv = a.wqkv != nullptr;
  if (P.G == 2) {
    P.nr[0] = ((a.rows + 1) / 2 + 7) / 8 * 8;
    if (P.nr[0] > 64) P.nr[0] = 64;
    P.nr[1] = a.rows - P.nr[0];
    VLG_REQUIRE(P.nr[1] > 0 && P.nr[1] <= 64, "layer_tail2: bad row split %d + %d", P.nr[0], P.nr[1]);
  } else {
    P.nr[0] = a.rows; P.nr[1] = 0;
  }
  P.row0[0] = 0; P.row0[1] = P.nr[0];
  const uint32_t a_rows = (uint32_t)((P.nr[0] + 7) / 8 * 8);
  P.a_bytes = a_rows * 128u;
  const LtGemm* gs[2] = {&P.g_wo, &P.g_w2};
  for (const LtGemm* g : gs)
    VLG_REQUIRE((size_t)g->S * a.rows * g->N * sizeof(float) <= a.part_bytes, "layer_tail2: scratch too small (%zu bytes)", a.part_bytes);
  P.part = a.part;
  P.h = static_cast<__nv_bfloat16*>(a.h);
  P.xn = static_cast<__nv_bfloat16*>(a.xn);
  P.norm1_w = static_cast<const __nv_bfloat16*>(a.norm1_w);
  P.norm2_w = static_cast<const __nv_bfloat16*>(a.norm2_w);
  P.eps = a.eps;
  P.ffn.out = a.ffn; P.ffn.ld_out = a.hidden;
  P.rope = a.rope;
  P.cnt = a.bar;
  P.err = a.err;
  static const int dbg = [] { const char* e = getenv("VLG_LT_DEBUG"); return e ? atoi(e) : 0; }();
  P.dbg = dbg;
  static const int fast_silu = [] { const char* e = getenv("VLG_LT_FASTSILU"); return e ? atoi(e) : 1; }();
  P.fast_silu = fast_silu;

  T2Maps tm{};
  int rc;
  if ((rc = make_tmap_bf16_2d(&tm.xn, a.xn, (uint64_t)a.rows, (uint64_t)a.dim, (uint64_t)a.dim * 2, a_rows, 64, true))) return rc;
  if (P.has_front) {
    if ((rc = make_tmap_bf16_2d(&tm.attn, a.attn, (uint64_t)a.rows, (uint64_t)a.dim, (uint64_t)a.dim * 2, a_rows, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.ffn, a.ffn, (uint64_t)a.rows, (uint64_t)a.hidden, (uint64_t)a.hidden * 2, a_rows, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.wo, a.wo, (uint64_t)a.dim, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_wo.BN, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.w13, a.w13, (uint64_t)2 * a.hidden, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_w13.BN, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tm.w2, a.w2, (uint64_t)a.dim, (uint64_t)a.hidden, (uint64_t)a.hidden * 2, P.g_w2.BN, 64, true))) return rc;
  }
  if (P.has_qkv)
    if ((rc = make_tmap_bf16_2d(&tm.wqkv, a.wqkv, (uint64_t)3 * a.dim, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_qkv.BN, 64, true))) return rc;

  static PerDeviceOnce attr_once;
  VLG_CUDA(attr_once([&] { return cudaFuncSetAttribute(layer_tail2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t2_smem_bytes()); }));
  VLG_CUDA(launch_kernel(layer_tail2_kernel, dim3(num_sms()), dim3(kT2Threads), (size_t)t2_smem_bytes(), st, tm, P));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg

// instrumentation: stamps (ns), [160 CTAs][64 slots], of the last full layer_tail2 launch run with VLG_LT_DEBUG=1
extern "C" int vlg_debug_lt2_timestamps(unsigned long long* out_host) {
  VLG_CUDA(cudaMemcpyFromSymbol(out_host, vlg::g_lt2_ts, sizeof(unsigned long long) * 160 * 64));
  return VLG_OK;
}
