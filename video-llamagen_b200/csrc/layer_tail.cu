// Everything between two attentions of the bf16 decode step as ONE persistent kernel:
//
//   wo(attn) + residual -> RMSNorm -> w1|w3 -> SwiGLU -> w2 + residual -> RMSNorm -> wqkv(next layer) -> RoPE + KV append
//   (gpt.py:255-259 TransformerBlock, :143-148 RMSNorm, :151-167 FeedForward, :207-227 Attention up to the cache write)
//
// Why: at R <= 128 rows these GEMMs move 2-11 MB of weights each -- ~1 us of HBM time -- and as separate launches each of
// them paid a kernel boundary, a cold TMA round trip and, worst, the whole [R x K] activation slab through every SM's
// ~64 B/clk L2 port (profiles/README.md).  Here one CTA per SM stays resident and the step is a sequence of phases
// separated by grid barriers (one L2 atomic + one polled load, all CTAs co-resident by construction):
//
//   * GEMM phases: a unit = (BN-wide column tile, K slice), at most one unit per CTA, chosen on the host (lt_pick) from
//     the bytes a CTA pulls through its L2 port.  Operands arrive through TMA (128B swizzle) in a 6-stage ring, one
//     elected thread issues tcgen05.mma (M = 128 rows, N = BN, K = 16) into a TMEM accumulator, eight warps (two per TMEM
//     lane quadrant) read it back with tcgen05.ld.
//       - wo and w2 are split-K (their residual + RMSNorm row phase follows anyway): every SM ingests only its
//         [128 x K/S] activation slice and stores the fp32 PARTIAL tile to an L2-resident scratch with 256-bit stores;
//       - w1|w3 and wqkv own all of K and apply SwiGLU / RoPE + cache append straight from TMEM with 32-byte stores
//         (split-K here would cost two more barriers and 13 MB more partial traffic than it saves in ingest).
//     The weight tiles of the NEXT GEMM phase are requested as soon as the ring has room, i.e. while this phase's
//     epilogue, the barrier and the row phase in between run: weights never wait on HBM latency.
//   * Row phase (one row per CTA): fixed-order reduction of the S partials (deterministic), residual add and RMSNorm with
//     the reference's rounding points (epilogue.cuh / elementwise.cu).
//
// Only the tensor-core tile is re-shaped; no arithmetic of the reference is re-associated except the fp32 K-sum order.
#include "lt_common.cuh"

namespace vlg {

#ifndef VLG_LT_THREADS
#define VLG_LT_THREADS 384
#endif
#ifndef VLG_LT_STAGES
#define VLG_LT_STAGES 6
#endif
constexpr int kLtThreads = VLG_LT_THREADS;            // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: TMEM readers (two per lane quadrant); all: row phases
constexpr int kLtStages = VLG_LT_STAGES;
constexpr int kLtReaders = (kLtThreads / 32 - 2) >= 8 ? 8 : 4;   // TMEM reader warps: two per lane quadrant when the block has them
constexpr int kLtHalves = kLtReaders / 4;
constexpr int kLtABytes = 128 * 64 * 2;    // activation tile: 128 rows x 64 k
// Ring layouts (template parameters of the kernel): a stage = activation tile + weight tile slot.  Shapes whose column tiles
// are all <= 64 wide use 24 KB stages and a 9-deep ring, the others 32 KB stages and 6: the ingest-bound phases want as
// many bytes in flight as shared memory holds (4 / 5 / 6 stages of 32 KB: 831 / 809 / 800 ms per cfg-2 generate).
constexpr int kLtStageBytesWide = 2 * kLtABytes;          // weight tile of up to 128 rows x 64 k
constexpr int kLtStageBytesNarrow = kLtABytes + 8192;     // weight tile of up to 64 rows x 64 k
constexpr int kLtStagesNarrow = 9;
// pair mode (see lt_gemm_phase): a stage = this CTA's 64 activation rows + its half (<= 64 rows) of the pair's weight tile
constexpr int kLtPairABytes = 64 * 128;
constexpr int kLtStageBytesPair = kLtPairABytes + 8192;
constexpr int kLtStagesPair = 13;
constexpr int kLtTmemCols = 256;          // two accumulators (one per MMA-issuing warp), 128 columns apart
constexpr int kLtIssuer2 = 2 + kLtReaders;   // second MMA-issuing warp: the first warp behind the TMEM readers
constexpr int lt_smem_bytes(int stages, int stage_bytes) { return stages * stage_bytes + 1024 /*align slack*/ + 512 /*barriers*/; }

struct LtParams {
  int rows, dim, hidden;
  int has_front;   // wo .. w2 phases present (false for the launch in front of layer 0)
  int has_qkv;     // wqkv + RoPE phases present (false after the last layer)
  LtGemm g_wo, g_w13, g_w2, g_qkv;
  float* part;               // split-K partials [S][rows][N] fp32, reused by every GEMM phase
  __nv_bfloat16* h;          // residual stream [rows, dim], updated in place
  __nv_bfloat16* xn;         // normed hidden   [rows, dim]   (A operand of w13 / wqkv / the head)
  const __nv_bfloat16* norm1_w;   // ffn_norm of this layer
  const __nv_bfloat16* norm2_w;   // attention_norm of the next layer, or the final norm
  float eps;
  EpiParams ffn;             // SwiGLU destination (out, ld_out)
  EpiParams rope;            // RoPE + cache append parameters
  unsigned* bar;             // [0] grid-barrier arrivals, [1] exit arrivals; both zero between launches
  volatile unsigned* err;    // mapped host word: set when a grid barrier gave up waiting (see lt_grid_barrier)
  int dbg;
  uint32_t a_bytes;          // bytes of one activation tile as TMA delivers it: round8(rows) x 128 (the MMA reads 128 rows; the
                             // rows beyond are whatever the stage held before -- they only feed accumulator lanes nobody reads)
  int fast_silu;             // SwiGLU epilogue: ex2.approx / rcp.approx instead of expf + IEEE division (A/B: VLG_LT_FASTSILU)
  // L2 prefetch of what the NEXT kernel streams from HBM (see lt_pf_slice): the K/V prefix of the next layer's attention,
  // and/or one flat region (the head GEMM's weights after the last layer)
  const uint8_t* pf_k;       // next layer's K cache [n_rh][S_max][128 B]
  const uint8_t* pf_v;
  int pf_n_rh;               // rows * heads
  long long pf_stride;       // bytes between (row, head) spans = S_max * 128
  const int* pf_pos_ptr;     // device cache position (rows [0, pos) of every span are complete), or null -> pf_pos_imm
  int pf_pos_imm;
  long long pf_budget;       // bytes of K + V to request per launch (0 = off)
  const uint8_t* pf_flat;    // flat region, pf_flat_bytes long
  long long pf_flat_bytes;
};

// phase timestamps (globaltimer ns) of every CTA of the last full (front + qkv) launch, VLG_LT_DEBUG=1: instrumentation
// for tools/lt_phases.py.  Slots 0-15: thread 0 (producer / barrier view); 16-31: first TMEM-reader thread (accumulator view).
__device__ unsigned long long g_lt_ts[160 * 32];
// EXTRA (template parameter of the kernel and of lt_gemm_phase): the instrumented build of the kernel -- phase timestamps, the
// optional L2 prefetch slices, the exact-SiLU switch.  The kernel the decode loop launches has none of it COMPILED IN: every
// phase runs once per launch, so its instructions are fetched cold, and those three features were 115 of the kernel's 194 KB
// of SASS (ncu: "no instruction" was the second-largest warp-state bucket).  79 KB: 778 -> 743 ms per cfg-2 generate.
#define LT_TS_AT(i)                                                                                   \
  do {                                                                                                \
    if constexpr (EXTRA) {                                                                            \
      if (P.dbg && P.has_front && P.has_qkv && blockIdx.x < 160) {                                    \
        unsigned long long t_;                                                                        \
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                         \
        g_lt_ts[blockIdx.x * 32 + (i)] = t_;                                                          \
      }                                                                                               \
    }                                                                                                 \
  } while (0)
#define LT_TS(i) do { if (threadIdx.x == 0) LT_TS_AT(i); } while (0)

// All CTAs of the grid are co-resident (grid <= SM count, one CTA per SM): arrive with a release add, poll with acquire.
// Co-residency holds whenever this context has the GPU's SMs to itself; if some other resident work keeps CTAs of this grid
// from being scheduled, the wait is bounded (2 s of globaltimer) and ends in an error word in mapped host memory that the
// host turns into VLG_ERR_STATE on its next call -- never in a trap that would poison the CUDA context.  Once the word is
// set every later barrier of every CTA falls through at once, so a contended launch drains in seconds, not minutes.
// Warp 0 (the TMA producer) is not part of the arrival: at the end of a GEMM phase it is still requesting the NEXT phase's
// weight tiles (waiting for ring slots, ~13 TMA issues), and with a CTA-wide __syncthreads in front of the arrival one late
// producer held back the whole grid (phase timeline: arrivals up to 1 us after the last epilogue store).  The workers (warps
// 1..) synchronise among themselves, one of them arrives and polls, and the producer joins at the closing __syncthreads.
constexpr int kLtWorkers = kLtThreads - 32;
__device__ __forceinline__ void lt_workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kLtWorkers) : "memory"); }

// the bounded-wait bookkeeping of a barrier poll loop, out of line (six barriers are inlined into the kernel)
static __device__ __noinline__ bool lt_barrier_gave_up(unsigned long long& t0, volatile unsigned* err) {
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  if (t0 == 0) t0 = now;
  if (err != nullptr && *err != 0) return true;                       // somebody already gave up
  if (now - t0 > 2000000000ull) {
    if (err != nullptr) { *err = 1u; __threadfence_system(); }
    return true;
  }
  return false;
}

__device__ __forceinline__ void lt_grid_barrier(unsigned* ctr, unsigned& target, volatile unsigned* err) {
  target += gridDim.x;
  tc_fence_before();
  if (threadIdx.x >= 32) {
    lt_workers_sync();                      // every global store of this CTA's phase is issued (the producer makes none)
    if (threadIdx.x == 32) {
      // release: cumulative over the workers' stores ordered before it by the bar.sync above (tools/micro/bar_bench.cu: the
      // extra fences of the cooperative-groups recipe cost 0.25 us per barrier and add nothing to this pattern;
      // tools/micro/bar_bench2.cu: moving the pollers off the counter's line onto flag lines is slower, 1.8-2.0 vs 1.25 us)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      unsigned seen, spins = 0;
      unsigned long long t0 = 0;
      for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        if (seen >= target) break;
        if ((++spins & 1023u) == 0 && lt_barrier_gave_up(t0, err)) break;
      }
    }
  }
  __syncthreads();
  tc_fence_after();
}

__device__ __forceinline__ void lt_add_acc16(uint32_t taddr, float* v) {
  float u[16];
  tmem_ld16(taddr, u);
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __fadd_rn(v[i], u[i]);
}

template <int ST> struct LtSmem {
  uint64_t full[ST], empty[ST];
  uint64_t acc_ready;
  uint32_t tmem_slot;
  float red[kLtThreads / 32];
};

// ---------------------------------------------------------------------------------------------------- L2 prefetch
// This kernel is latency-bound (grid barriers, dependent phases) and keeps HBM busy for ~4 of its ~25 us, while the attention
// kernel that follows is HBM-bound.  So the last warp of every CTA spends the idle HBM time pulling the leading part of the
// next layer's K/V prefix into L2 (cp.async.bulk.prefetch.L2: fire-and-forget, no shared memory, no completion to wait
// for): the attention kernel then finds a mix of L2 hits and HBM misses in its TMA ring and streams at the sum of both
// rates.  Rows [0, pos) of the next layer's cache were written in earlier decode steps, so nothing here is ordered against
// this launch's own writes (row `pos` is written by the wqkv epilogue below and is not requested).  One slice per phase
// paces the requests over the whole launch so that they do not queue in front of this kernel's own operand loads.
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void lt_pf_slice(const LtParams& P, int slice, int n_slices) {
  if (threadIdx.x < kLtThreads - 32) return;            // the last warp: it owns no GEMM role
  const int lane = threadIdx.x & 31;
  constexpr long long kSub = 2048;
  if (P.pf_budget > 0 && P.pf_k != nullptr) {
    const int pos = P.pf_pos_ptr ? *P.pf_pos_ptr : P.pf_pos_imm;
    const long long span = (long long)pos * 128;
    const long long n_span = 2LL * P.pf_n_rh;
    if (span > 0 && n_span > 0) {
      long long per_span = n_span * span <= P.pf_budget ? span : ((P.pf_budget / n_span) & ~127LL);
      const long long piece = ((per_span + n_slices - 1) / n_slices + 127) & ~127LL;
      const long long off = (long long)slice * piece;
      if (off < per_span) {
        const long long len = per_span - off < piece ? per_span - off : piece;
        const long long n_sub = (len + kSub - 1) / kSub;
        const long long items = n_span * n_sub;
        for (long long w = (long long)blockIdx.x * 32 + lane; w < items; w += 32LL * gridDim.x) {
          const long long s = w / n_sub, c = w - s * n_sub;
          const uint8_t* base = (s < P.pf_n_rh ? P.pf_k + s * P.pf_stride : P.pf_v + (s - P.pf_n_rh) * P.pf_stride) + off + c * kSub;
          const long long b = len - c * kSub < kSub ? len - c * kSub : kSub;
          bulk_prefetch_l2(base, (uint32_t)b);
        }
      }
    }
  }
  if (P.pf_flat != nullptr && P.pf_flat_bytes > 0) {
    const long long piece = ((P.pf_flat_bytes + n_slices - 1) / n_slices + kSub - 1) / kSub * kSub;
    const long long off = (long long)slice * piece;
    if (off < P.pf_flat_bytes) {
      const long long len = P.pf_flat_bytes - off < piece ? P.pf_flat_bytes - off : piece;
      const long long items = (len + kSub - 1) / kSub;
      for (long long w = (long long)blockIdx.x * 32 + lane; w < items; w += 32LL * gridDim.x) {
        const long long b = len - w * kSub < kSub ? len - w * kSub : kSub;
        bulk_prefetch_l2(P.pf_flat + off + w * kSub, (uint32_t)(b & ~15LL));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------- GEMM phase
// g: k-blocks this CTA has pushed through the ring so far (producer and MMA issuer walk the same sequence);
// acc_uses: GEMM phases in which this CTA owned a unit (parity of acc_ready).
// Unit of this CTA in a GEMM phase.  PAIR: the two CTAs of a cluster own one unit; rank r loads activation rows [64 r, 64 r + 64)
// and rows [n0 + r BN/2, + BN/2) of the weight tile, the leader (rank 0) issues ONE tcgen05.mma.cta_group::2 per K step that
// reads both halves (pair_gemm.cu is the verified reference of the mechanics): the bytes through each SM's L2 port drop from
// (128 + BN) to (64 + BN/2) rows per k-block.  `full` barriers live in the leader and count the bytes of both CTAs.
template <bool PAIR> __device__ __forceinline__ int lt_unit_id() { return PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x; }

template <int ST, int SB, bool PAIR>
__device__ __forceinline__ void lt_prefetch_b(const LtGemm& G, const TensorMap* tmB, uint8_t* ring, LtSmem<ST>& sm, uint32_t g, uint32_t P_a_bytes) {
  const LtUnit v = lt_unit(G, lt_unit_id<PAIR>());
  if (!v.valid) return;
  const int pre = v.nkb < ST ? v.nkb : ST;
  if constexpr (PAIR) {
    const uint32_t rank = blockIdx.x & 1, tx = 2 * (kLtPairABytes + G.BN * 64);
    for (int kb = 0; kb < pre; ++kb) {
      const uint32_t s = (g + kb) % ST, it = (g + kb) / ST;
      mbar_wait(&sm.empty[s], (it & 1) ^ 1);
      if (rank == 0) mbar_expect_tx(&sm.full[s], tx);
      tma_load_2d_pair(ring + s * SB + kLtPairABytes, tmB, &sm.full[s], (v.kb0 + kb) * 64, v.n0 + (int)rank * (G.BN >> 1), kEvictFirst);
    }
  } else {
    const uint32_t tx = P_a_bytes + G.BN * 128;
    for (int kb = 0; kb < pre; ++kb) {
      const uint32_t s = (g + kb) % ST, it = (g + kb) / ST;
      mbar_wait(&sm.empty[s], (it & 1) ^ 1);
      mbar_expect_tx(&sm.full[s], tx);
      tma_load_2d_hint(ring + s * SB + kLtABytes, tmB, &sm.full[s], (v.kb0 + kb) * 64, v.n0, kEvictFirst);
    }
  }
}

// MODE: -1 = store the fp32 partial tile (split-K, reduced by the following row phase); EPI_SWIGLU / EPI_QKV_ROPE = the unit
// owns all of K and applies the elementwise tail straight from TMEM (lane == row).
template <int MODE, int ST, int SB, bool PAIR, bool EXTRA>
__device__ __forceinline__ void lt_gemm_phase(const LtGemm& G, const TensorMap* tmA, const TensorMap* tmB,
                                              const LtGemm* Gnext, const TensorMap* tmBnext, const LtParams& P,
                                              const EpiParams& E, uint8_t* ring, LtSmem<ST>& sm, uint32_t tmem_acc, uint32_t& g, uint32_t& acc_uses,
                                              int ts_slot) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const LtUnit u = lt_unit(G, lt_unit_id<PAIR>());
  const uint32_t nkb = u.valid ? (uint32_t)u.nkb : 0u;
  const uint32_t rank = PAIR ? (blockIdx.x & 1) : 0u;
  if (warp == 0) {
    if (elect_one()) {
      asm volatile("fence.proxy.async;" ::: "memory");     // other SMs' generic-proxy stores of A -> this SM's TMA reads
      if (u.valid) {
        const int pre = u.nkb < ST ? u.nkb : ST;      // their weight tiles + expect_tx were issued a phase ago
        if constexpr (PAIR) {
          for (int kb = 0; kb < pre; ++kb) {
            const uint32_t s = (g + kb) % ST;
            tma_load_2d_pair(ring + s * SB, tmA, &sm.full[s], (u.kb0 + kb) * 64, (int)rank * 64, kEvictLast);
          }
          const uint32_t tx = 2 * (kLtPairABytes + G.BN * 64);
          for (int kb = pre; kb < u.nkb; ++kb) {
            const uint32_t s = (g + kb) % ST, it = (g + kb) / ST;
            mbar_wait(&sm.empty[s], (it & 1) ^ 1);
            if (rank == 0) mbar_expect_tx(&sm.full[s], tx);
            uint8_t* a = ring + s * SB;
            tma_load_2d_pair(a, tmA, &sm.full[s], (u.kb0 + kb) * 64, (int)rank * 64, kEvictLast);
            tma_load_2d_pair(a + kLtPairABytes, tmB, &sm.full[s], (u.kb0 + kb) * 64, u.n0 + (int)rank * (G.BN >> 1), kEvictFirst);
          }
        } else {
          for (int kb = 0; kb < pre; ++kb) {
            const uint32_t s = (g + kb) % ST;
            tma_load_2d_hint(ring + s * SB, tmA, &sm.full[s], (u.kb0 + kb) * 64, 0, kEvictLast);
          }
          const uint32_t tx = P.a_bytes + G.BN * 128;
          for (int kb = pre; kb < u.nkb; ++kb) {
            const uint32_t s = (g + kb) % ST, it = (g + kb) / ST;
            mbar_wait(&sm.empty[s], (it & 1) ^ 1);
            mbar_expect_tx(&sm.full[s], tx);
            uint8_t* a = ring + s * SB;
            tma_load_2d_hint(a, tmA, &sm.full[s], (u.kb0 + kb) * 64, 0, kEvictLast);
            tma_load_2d_hint(a + kLtABytes, tmB, &sm.full[s], (u.kb0 + kb) * 64, u.n0, kEvictFirst);
          }
        }
      }
      if (Gnext != nullptr) lt_prefetch_b<ST, SB, PAIR>(*Gnext, tmBnext, ring, sm, g + nkb, P.a_bytes);
    }
    __syncwarp();
  } else if (warp == 1 || warp == kLtIssuer2) {
    // Two issuing threads, alternating k-blocks, each into its own accumulator (summed by the readers): one thread sustains a
    // tcgen05.mma per ~130 cycles whatever the tile shape (tools/micro/mma_issue_bench.cu), i.e. 0.26 us per k-block -- more
    // than the 0.18-0.22 us the SM's L2 port needs to deliver the k-block's operands.
    const int j = warp == 1 ? 0 : 1;
    if (u.valid && rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, G.BN);
      const uint32_t acc = tmem_acc + (PAIR ? 64u : 128u) * j;    // a pair's accumulator is BN/2 <= 64 columns wide
      for (int kb = j; kb < u.nkb; kb += 2) {
        const uint32_t s = (g + kb) % ST, it = (g + kb) / ST;
        mbar_wait(&sm.full[s], it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(ring + s * SB);
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + (PAIR ? kLtPairABytes : kLtABytes));
        if constexpr (PAIR) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_pair(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb >= 2 || k != 0));
          umma_commit_pair(&sm.empty[s]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb >= 2 || k != 0));
          umma_commit(&sm.empty[s]);
        }
      }
      if constexpr (PAIR) umma_commit_pair(&sm.acc_ready);
      else umma_commit(&sm.acc_ready);
    }
    __syncwarp();
  } else if (warp < 2 + kLtReaders) {
    if (u.valid) {
      // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the columns.  Every store below
      // covers whole 32-byte sectors: partial-sector writes cost the L2 several times their size.
      // PAIR: this CTA holds rows [64 rank, 64 rank + 64) in the 2x2 layout -- lanes [0, 64) carry columns [0, BN/2) of the
      // pair's tile, lanes [64, 128) columns [BN/2, BN) -- so a quadrant's two warps split BN/2 columns, not BN.
      const int quad = warp & 3, half = (warp - 2) >> 2;
      const int r = PAIR ? (int)rank * 64 + (quad & 1) * 32 + lane : quad * 32 + lane;
      const int ncol = PAIR ? (G.BN >> 1) : G.BN;                 // columns on this warp's lanes
      const int cbase = PAIR ? (quad >> 1) * (G.BN >> 1) : 0;     // tile column of TMEM column 0 on these lanes
      constexpr uint32_t kAcc2 = PAIR ? 64u : 128u;               // TMEM column distance of the second issuer's accumulator
      const uint32_t lane_base = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);
      const bool two = u.nkb > 1;            // the second issuer's accumulator holds the odd k-blocks
      // RoPE: the cache position and the first chunk's (cos, sin) pairs are two dependent global loads; request them while
      // the main loop still runs instead of after the accumulator is ready
      int rope_pos = 0;
      float2 rope_cs[8];
      if constexpr (MODE == EPI_QKV_ROPE) {
        rope_pos = E.pos_ptr ? *E.pos_ptr : E.pos_imm;
        const int n = u.n0 + cbase + half * 16, sec = n / E.dim, cc = n - sec * E.dim, dd = cc % E.D;
        if (sec < 2) {
          const float2* tab = reinterpret_cast<const float2*>(E.rope) + (size_t)rope_pos * (E.D >> 1) + (dd >> 1);
#pragma unroll
          for (int i = 0; i < 8; ++i) rope_cs[i] = tab[i];
        }
      }
      mbar_wait(&sm.acc_ready, acc_uses & 1);
      tc_fence_after();
      if (threadIdx.x == 64) LT_TS_AT(ts_slot);
      if constexpr (MODE < 0) {
        float* dst = P.part + ((size_t)u.split * P.rows + r) * G.N + u.n0 + cbase;
        for (int c = half * 16; c < ncol; c += 16 * kLtHalves) {
          float v[16];
          tmem_ld16(lane_base + c, v);
          if (two) lt_add_acc16(lane_base + kAcc2 + c, v);
          if (r < P.rows) {
            stg256(dst + c, v);
            stg256(dst + c + 8, v + 8);
          }
        }
      } else if constexpr (MODE == EPI_SWIGLU) {      // gpt.py:167 -- w1/w3 rows interleaved: columns (2j, 2j+1) = (w1 x, w3 x)_j
        for (int c = half * 32; c < ncol; c += 32 * kLtHalves) {
          float v[32];
          tmem_ld16(lane_base + c, v);
          tmem_ld16(lane_base + c + 16, v + 16);
          if (threadIdx.x == 64) LT_TS_AT(24);
          if (two) { lt_add_acc16(lane_base + kAcc2 + c, v); lt_add_acc16(lane_base + kAcc2 + c + 16, v + 16); }
          if (threadIdx.x == 64) LT_TS_AT(25);
          if (r < P.rows) {
            float o[16];
            // one branch around two straight-line loops: with the choice made per element ptxas keeps the 16 independent
            // MUFU chains in sequence (measured: 2.0 us for this block, most of the phase's epilogue)
            if (!EXTRA || P.fast_silu) {       // 2-3 ulp of fp32 before the bf16 rounding of silu (2^-9 relative)
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                float a0 = v[2 * i], b0 = v[2 * i + 1], a1 = v[2 * i + 2], b1 = v[2 * i + 3];
                rnd2_bf16(a0, b0);
                rnd2_bf16(a1, b1);
                float s0 = __fdividef(a0, 1.0f + __expf(-a0)), s1 = __fdividef(a1, 1.0f + __expf(-a1));
                rnd2_bf16(s0, s1);
                o[i] = __fmul_rn(s0, b0);
                o[i + 1] = __fmul_rn(s1, b1);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float a = rnd<__nv_bfloat16>(v[2 * i]), b = rnd<__nv_bfloat16>(v[2 * i + 1]);
                o[i] = __fmul_rn(rnd<__nv_bfloat16>(silu_f(a)), b);
              }
            }
            if (threadIdx.x == 64) LT_TS_AT(26);
            stg256_bf16x16(static_cast<__nv_bfloat16*>(E.out) + (size_t)r * E.ld_out + ((u.n0 + cbase + c) >> 1), o);
          }
        }
      } else {                                         // EPI_QKV_ROPE on 16 columns (a quarter head): gpt.py:215-227, 423-433, 177-185
        const int pos = rope_pos;
        for (int c = half * 16; c < ncol; c += 16 * kLtHalves) {
          float v[16];
          tmem_ld16(lane_base + c, v);
          if (two) lt_add_acc16(lane_base + kAcc2 + c, v);
          if (r < P.rows) {
            const int n = u.n0 + cbase + c, sec = n / E.dim, cc = n - sec * E.dim, head = cc / E.D, dd = cc - head * E.D;
            float o[16];
            if (sec < 2) {
              const float2* tab = reinterpret_cast<const float2*>(E.rope) + (size_t)pos * (E.D >> 1) + (dd >> 1);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 cs = c == half * 16 ? rope_cs[i] : tab[i];
                float x0 = v[2 * i], x1 = v[2 * i + 1];
                rnd2_bf16(x0, x1);
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
      if (threadIdx.x == 64) LT_TS_AT(ts_slot + 1);
    }
  }
  g += nkb;
  acc_uses += u.valid ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------------------- row phases
// h <- T(h + T(sum of partials))  (when G != null), then xn <- T(T(h * rsqrt(mean(h^2) + eps)) * w): one row per CTA
template <int ST>
__device__ __forceinline__ void lt_norm_phase(const LtGemm* G, const __nv_bfloat16* w, const LtParams& P, LtSmem<ST>& sm) {
  if (threadIdx.x < 32) return;            // the producer warp keeps requesting weight tiles; the workers own the row phases
  const int nvec = P.dim / 8, t = threadIdx.x - 32;
  const bool active = t < nvec;
  float g8[8];
  if (active) load8<__nv_bfloat16>(w + t * 8, g8);
  for (int r = blockIdx.x; r < P.rows; r += gridDim.x) {
    float hn[8];
    float ss = 0.f;
    if (active) {
      __nv_bfloat16* hp = P.h + (size_t)r * P.dim + t * 8;
      const uint4 hu = ldg128_cg(hp);
      hn[0] = bf16lo(hu.x); hn[1] = bf16hi(hu.x); hn[2] = bf16lo(hu.y); hn[3] = bf16hi(hu.y);
      hn[4] = bf16lo(hu.z); hn[5] = bf16hi(hu.z); hn[6] = bf16lo(hu.w); hn[7] = bf16hi(hu.w);
      if (G != nullptr) {
        float acc[8];
        lt_reduce8(P.part, G->S, (size_t)P.rows * G->N, (size_t)r * G->N + t * 8, acc);
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          rnd2_bf16(acc[i], acc[i + 1]);
          hn[i] = __fadd_rn(hn[i], acc[i]);
          hn[i + 1] = __fadd_rn(hn[i + 1], acc[i + 1]);
          rnd2_bf16(hn[i], hn[i + 1]);
        }
        store8<__nv_bfloat16>(hp, hn);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = __fadd_rn(ss, __fmul_rn(hn[i], hn[i]));
    }
    ss = warp_sum(ss);
    lt_workers_sync();                                 // previous row's readers of sm.red are done
    if ((t & 31) == 0) sm.red[t >> 5] = ss;
    lt_workers_sync();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kLtWorkers / 32; ++i) tot += sm.red[i];
    const float rs = rsqrtf(__fdiv_rn(tot, (float)P.dim) + P.eps);
    if (active) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        float y0 = __fmul_rn(hn[i], rs), y1 = __fmul_rn(hn[i + 1], rs);
        rnd2_bf16(y0, y1);
        o[i] = __fmul_rn(y0, g8[i]);
        o[i + 1] = __fmul_rn(y1, g8[i + 1]);
      }
      store8<__nv_bfloat16>(P.xn + (size_t)r * P.dim + t * 8, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------- the kernel
template <int ST, int SB, bool PAIR, bool EXTRA>
__global__ void __launch_bounds__(kLtThreads, 1)
layer_tail_kernel(const __grid_constant__ TensorMap tmAttn, const __grid_constant__ TensorMap tmXn,
                  const __grid_constant__ TensorMap tmFfn, const __grid_constant__ TensorMap tmWo,
                  const __grid_constant__ TensorMap tmW13, const __grid_constant__ TensorMap tmW2,
                  const __grid_constant__ TensorMap tmWqkv, const LtParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  LtSmem<ST>& sm = *reinterpret_cast<LtSmem<ST>*>(ring + ST * SB);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    if (P.has_front) { tma_prefetch_desc(&tmAttn); tma_prefetch_desc(&tmFfn); tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW13); tma_prefetch_desc(&tmW2); }
    tma_prefetch_desc(&tmXn);
    if (P.has_qkv) tma_prefetch_desc(&tmWqkv);
    for (int s = 0; s < ST; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.acc_ready, 2);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair<128>(&sm.tmem_slot);     // two accumulators of <= 64 columns
    else tmem_alloc<kLtTmemCols>(&sm.tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();                         // the partner's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = sm.tmem_slot;

  LT_TS(0);
  pdl_launch_dependents();
  uint32_t g = 0, acc_uses = 0;
  unsigned bar_target = 0;
  // the first GEMM's weight tiles do not depend on the previous kernel: request them before waiting for it
  if (warp == 0) {
    if (elect_one()) {
      if (P.has_front) lt_prefetch_b<ST, SB, PAIR>(P.g_wo, &tmWo, ring, sm, 0, P.a_bytes);
      else if (P.has_qkv) lt_prefetch_b<ST, SB, PAIR>(P.g_qkv, &tmWqkv, ring, sm, 0, P.a_bytes);
    }
    __syncwarp();
  }
  pdl_wait();
  LT_TS(1);

  // (Measured: the phases as a loop over a small state machine -- wo / w2 and the two row phases sharing one copy of their code,
  // one barrier call site; 77 -> 65 KB of SASS -- is slower, 847 against 743 ms: operands picked by reference at run time move
  // the parameter block to local memory.  With the operands copied by value (a two-trip loop around wo / w2, 59 KB, no stack
  // frame) it is exactly as fast as this straight-line code, 740.8 against 740.3 ms: below ~77 KB the code size no longer
  // matters.  Straight-line code it stays.)
  [[maybe_unused]] const int pf_slices = (P.has_front ? 5 : 1) + (P.has_qkv ? 1 : 0);
  [[maybe_unused]] int pf_i = 0;
  if (P.has_front) {
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_gemm_phase<-1, ST, SB, PAIR, EXTRA>(P.g_wo, &tmAttn, &tmWo, &P.g_w13, &tmW13, P, P.ffn, ring, sm, tmem_acc, g, acc_uses, 16);
    LT_TS(2);
    lt_grid_barrier(P.bar, bar_target, P.err);
    LT_TS(3);
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_norm_phase(&P.g_wo, P.norm1_w, P, sm);
    LT_TS(4);
    lt_grid_barrier(P.bar, bar_target, P.err);
    LT_TS(5);
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_gemm_phase<EPI_SWIGLU, ST, SB, PAIR, EXTRA>(P.g_w13, &tmXn, &tmW13, &P.g_w2, &tmW2, P, P.ffn, ring, sm, tmem_acc, g, acc_uses, 18);
    LT_TS(6);
    lt_grid_barrier(P.bar, bar_target, P.err);
    LT_TS(7);
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_gemm_phase<-1, ST, SB, PAIR, EXTRA>(P.g_w2, &tmFfn, &tmW2, P.has_qkv ? &P.g_qkv : nullptr, &tmWqkv, P, P.ffn, ring, sm, tmem_acc, g, acc_uses, 20);
    LT_TS(8);
    lt_grid_barrier(P.bar, bar_target, P.err);
    LT_TS(9);
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_norm_phase(&P.g_w2, P.norm2_w, P, sm);
    LT_TS(10);
  } else {
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_norm_phase(nullptr, P.norm2_w, P, sm);
  }
  if (P.has_qkv) {
    lt_grid_barrier(P.bar, bar_target, P.err);
    LT_TS(11);
    if constexpr (EXTRA) lt_pf_slice(P, pf_i++, pf_slices);
    lt_gemm_phase<EPI_QKV_ROPE, ST, SB, PAIR, EXTRA>(P.g_qkv, &tmXn, &tmWqkv, nullptr, nullptr, P, P.rope, ring, sm, tmem_acc, g, acc_uses, 22);
    LT_TS(12);
  }

  // leave the counters at zero for the next launch: the last CTA out has seen everyone past the final barrier
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(P.bar + 1, 1u);
    if (old == gridDim.x - 1) {
      P.bar[0] = 0;
      P.bar[1] = 0;
      __threadfence();
    }
  }
  if constexpr (PAIR) {
    cluster_sync_all();                    // the leader's MMAs read the partner's shared memory and write its tensor memory
    if (warp == 1) tmem_dealloc_pair<128>(tmem_acc);
  } else {
    if (warp == 1) tmem_dealloc<kLtTmemCols>(tmem_acc);
  }
}

// ---------------------------------------------------------------------------------------------------- host
// Tile shape and split count per GEMM: bytes a CTA pulls through its L2 port (activation + weight slices), pushes out
// (partial tile) and the row phase pulls back, with at most one unit per CTA.
// Split-K units write fp32 partials: the L2 write path sustains only ~4.3 TB/s over the chip (tools/micro/bar_bench.cu), so
// a stored byte is costed 4x a loaded one.  Units that own all of K (split == false) need room for one tile per CTA.
bool lt_pick(int N, int K, int sms, bool split, LtGemm* out, int max_bn) {
  if (K % 64 != 0 || N % 64 != 0) return false;
  const int nkb = K / 64;
  // split count cap: the row phase that reduces the partials keeps 8 loads per thread in flight per L2 round trip
  static const int max_s = [] { const char* e = getenv("VLG_LT_MAXS"); const int v = e ? atoi(e) : kLtMaxSplits; return v < 1 ? 1 : v > kLtMaxSplits ? kLtMaxSplits : v; }();
  double best = 1e30;
  bool found = false;
  const int bns[3] = {32, 64, 128};
  for (int bn : bns) {
    if (N % bn || bn > max_bn) continue;
    const int tiles = N / bn;
    if (tiles > sms) continue;
    int s0 = split ? sms / tiles : 1;
    if (s0 > nkb) s0 = nkb;
    if (s0 > max_s) s0 = max_s;
    const int kb_per = (nkb + s0 - 1) / s0;
    const int S = (nkb + kb_per - 1) / kb_per;
    // costed at the full 128-row tile so that the choice (and the scratch size) does not depend on the batch
    double cost = (double)kb_per * (128 + bn) * 128.0;
    if (split) cost += 4.0 * 128.0 * bn * 4.0 + (double)S * 128 * N * 4.0 / sms;
    if (cost < best) {
      best = cost;
      found = true;
      out->N = N; out->K = K; out->BN = bn; out->S = S; out->n_tiles = tiles; out->kb_per = kb_per;
    }
  }
  return found;
}

// The same for CTA pairs: `pairs` units, column tile BN of the PAIR (each CTA holds BN/2 weight rows and 64 activation rows).
// min_bn: 128 where the epilogue needs 32 columns per reader warp (SwiGLU), 64 otherwise.
static bool lt_pick_pair(int N, int K, int pairs, bool split, int min_bn, LtGemm* out) {
  if (K % 64 != 0) return false;
  const int nkb = K / 64;
  static const int max_s = [] { const char* e = getenv("VLG_LT_MAXS"); const int v = e ? atoi(e) : kLtMaxSplits; return v < 1 ? 1 : v > kLtMaxSplits ? kLtMaxSplits : v; }();
  double best = 1e30;
  bool found = false;
  for (int bn : {64, 128}) {
    if (bn < min_bn || N % bn) continue;
    const int tiles = N / bn;
    if (tiles > pairs) continue;
    int s0 = split ? pairs / tiles : 1;
    if (s0 > nkb) s0 = nkb;
    if (s0 > max_s) s0 = max_s;
    const int kb_per = (nkb + s0 - 1) / s0;
    const int S = (nkb + kb_per - 1) / kb_per;
    double cost = (double)kb_per * (64 + bn / 2) * 128.0;                 // bytes through one SM's port
    if (split) cost += 4.0 * 64.0 * bn * 4.0 + (double)S * 128 * N * 4.0 / (2 * pairs);
    if (cost < best) {
      best = cost;
      found = true;
      out->N = N; out->K = K; out->BN = bn; out->S = S; out->n_tiles = tiles; out->kb_per = kb_per;
    }
  }
  return found;
}

// The grid barriers need every cluster resident at once: ask the driver how many 2-CTA clusters of the pair kernel the device
// holds (GPC shapes or a partitioned device can strand SMs) and use pairs only when that covers the whole grid.
static bool lt_pairs_fit(int sms) {
  static int cached[kMaxDevices] = {};                      // 0 unknown, 1 yes, -1 no
  int& c = cached[current_device()];
  if (c == 0) {
    auto kern = layer_tail_kernel<kLtStagesPair, kLtStageBytesPair, true, false>;
    const int smem = lt_smem_bytes(kLtStagesPair, kLtStageBytesPair);
    int n = 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms); cfg.blockDim = dim3(kLtThreads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const bool ok = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess &&
                    cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && 2 * n >= sms;
    if (!ok) cudaGetLastError();
    c = ok ? 1 : -1;
  }
  return c > 0;
}

// CTA pairs pay off when the activation slab is what fills the port: more than 64 rows.  VLG_LT_PAIR=0 keeps single CTAs (A/B).
static bool lt_pair_plan(int rows, int dim, int hidden, int sms, LtGemm* wo, LtGemm* w13, LtGemm* w2, LtGemm* qkv) {
  static const bool off = [] { const char* e = getenv("VLG_LT_PAIR"); return e && e[0] == '0'; }();
  if (off || rows <= 64 || sms < 2 || (sms & 1) || !lt_pairs_fit(sms)) return false;
  const int pairs = sms / 2;
  return lt_pick_pair(dim, dim, pairs, true, 64, wo) && lt_pick_pair(2 * hidden, dim, pairs, false, 128, w13) &&
         lt_pick_pair(dim, hidden, pairs, true, 64, w2) && lt_pick_pair(3 * dim, dim, pairs, false, 64, qkv);
}

bool layer_tail_supported(int rows, int dim, int hidden) {
  static const bool off = [] { const char* e = getenv("VLG_LT"); return e && e[0] == '0'; }();
  if (off) return false;
  if (rows <= 0 || rows > 128 || dim % 64 != 0 || hidden % 64 != 0 || dim / 8 > kLtThreads - 32) return false;
  const int sms = num_sms();
  LtGemm t;
  return lt_pick(dim, dim, sms, true, &t) && lt_pick(2 * hidden, dim, sms, false, &t) && lt_pick(dim, hidden, sms, true, &t) &&
         lt_pick(3 * dim, dim, sms, false, &t);
}

size_t layer_tail_scratch_bytes(int max_rows, int dim, int hidden) {
  const int rows = max_rows < 128 ? max_rows : 128;
  const int sms = num_sms();
  size_t worst = 0;
  const int shapes[2][2] = {{dim, dim}, {dim, hidden}};   // wo and w2 are the split-K GEMMs
  for (auto& sh : shapes)
    for (int max_bn : {64, 128}) {
      LtGemm t;
      if (!lt_pick(sh[0], sh[1], sms, true, &t, max_bn)) continue;
      const size_t b = (size_t)t.S * rows * t.N * sizeof(float);
      if (b > worst) worst = b;
    }
  LtGemm pw, p13, p2, pq;
  if (lt_pair_plan(128, dim, hidden, sms, &pw, &p13, &p2, &pq))
    for (const LtGemm* t : {&pw, &p2}) {
      const size_t b = (size_t)t->S * rows * t->N * sizeof(float);
      if (b > worst) worst = b;
    }
  return worst + 1024;
}

int launch_layer_tail(const LtLaunch& a, cudaStream_t st) {
  VLG_REQUIRE(layer_tail_supported(a.rows, a.dim, a.hidden), "layer_tail: unsupported shape rows=%d dim=%d hidden=%d", a.rows, a.dim, a.hidden);
  VLG_REQUIRE(a.wo || a.wqkv, "layer_tail: nothing to do");
  const int sms = num_sms();
  LtParams P{};
  P.rows = a.rows; P.dim = a.dim; P.hidden = a.hidden;
  P.has_front = a.wo != nullptr;
  P.has_qkv = a.wqkv != nullptr;
  const bool pair = lt_pair_plan(a.rows, a.dim, a.hidden, sms, &P.g_wo, &P.g_w13, &P.g_w2, &P.g_qkv);
  if (!pair) {
    lt_pick(a.dim, a.dim, sms, true, &P.g_wo);
    lt_pick(2 * a.hidden, a.dim, sms, false, &P.g_w13);
    lt_pick(a.dim, a.hidden, sms, true, &P.g_w2);
    lt_pick(3 * a.dim, a.dim, sms, false, &P.g_qkv);
  }
  const LtGemm* gs[2] = {&P.g_wo, &P.g_w2};
  for (const LtGemm* g : gs)
    VLG_REQUIRE((size_t)g->S * a.rows * g->N * sizeof(float) <= a.part_bytes, "layer_tail: scratch too small (%zu bytes)", a.part_bytes);
  P.part = a.part;
  P.h = static_cast<__nv_bfloat16*>(a.h);
  P.xn = static_cast<__nv_bfloat16*>(a.xn);
  P.norm1_w = static_cast<const __nv_bfloat16*>(a.norm1_w);
  P.norm2_w = static_cast<const __nv_bfloat16*>(a.norm2_w);
  P.eps = a.eps;
  P.ffn.out = a.ffn; P.ffn.ld_out = a.hidden;
  P.rope = a.rope;
  P.bar = a.bar;
  P.err = a.err;
  static const int dbg = [] { const char* e = getenv("VLG_LT_DEBUG"); return e ? atoi(e) : 0; }();
  P.dbg = dbg;
  static const int fast_silu = [] { const char* e = getenv("VLG_LT_FASTSILU"); return e ? atoi(e) : 1; }();
  P.fast_silu = fast_silu;
  // Activation tiles are loaded at their real height: a 128-row box over a 16-row tensor makes TMA zero-fill 14 KB of every
  // 16 KB tile, and the fill occupies the copy engine like real bytes do.  VLG_LT_FULLBOX=1 restores the padded box (A/B).
  static const bool full_box = [] { const char* e = getenv("VLG_LT_FULLBOX"); return e && e[0] == '1'; }();
  const uint32_t a_rows = pair ? 64u : full_box ? 128u : (uint32_t)((a.rows + 7) / 8 * 8);   // a pair's CTA loads 64 rows
  P.a_bytes = a_rows * 128u;
  // L2 prefetch budget per launch (MB of next-layer K + V).  OFF by default: measured on B200 (gpurun_out of round 2,
  // DESIGN.md 4.1) the attention kernel runs at the same speed from L2-resident K/V as from HBM (its shared-memory / issue
  // side bounds it at ~7.9 TB/s), while the fill traffic adds latency to every phase of this kernel: 819 / 846 / 902 / 954 ms
  // per cfg-2 generate at 0 / 48 / 80 / 112 MB.  VLG_KV_PREFETCH_MB=<n> re-enables it for A/B runs.
  static const long long pf_mb = [] { const char* e = getenv("VLG_KV_PREFETCH_MB"); return e ? atoll(e) : 0LL; }();
  if (pf_mb > 0 && a.pf_k != nullptr && a.pf_v != nullptr && a.pf_n_rh > 0) {
    P.pf_k = static_cast<const uint8_t*>(a.pf_k); P.pf_v = static_cast<const uint8_t*>(a.pf_v);
    P.pf_n_rh = a.pf_n_rh; P.pf_stride = a.pf_stride_bytes; P.pf_pos_ptr = a.pf_pos_ptr; P.pf_pos_imm = a.pf_pos_imm;
    P.pf_budget = pf_mb << 20;
  }
  if (pf_mb > 0 && a.pf_flat != nullptr) { P.pf_flat = static_cast<const uint8_t*>(a.pf_flat); P.pf_flat_bytes = (long long)a.pf_flat_bytes; }

  TensorMap tmAttn{}, tmXn{}, tmFfn{}, tmWo{}, tmW13{}, tmW2{}, tmWqkv{};
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmXn, a.xn, (uint64_t)a.rows, (uint64_t)a.dim, (uint64_t)a.dim * 2, a_rows, 64, true))) return rc;
  if (P.has_front) {
    if ((rc = make_tmap_bf16_2d(&tmAttn, a.attn, (uint64_t)a.rows, (uint64_t)a.dim, (uint64_t)a.dim * 2, a_rows, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmFfn, a.ffn, (uint64_t)a.rows, (uint64_t)a.hidden, (uint64_t)a.hidden * 2, a_rows, 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmWo, a.wo, (uint64_t)a.dim, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_wo.BN >> (pair ? 1 : 0), 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmW13, a.w13, (uint64_t)2 * a.hidden, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_w13.BN >> (pair ? 1 : 0), 64, true))) return rc;
    if ((rc = make_tmap_bf16_2d(&tmW2, a.w2, (uint64_t)a.dim, (uint64_t)a.hidden, (uint64_t)a.hidden * 2, P.g_w2.BN >> (pair ? 1 : 0), 64, true))) return rc;
  }
  if (P.has_qkv)
    if ((rc = make_tmap_bf16_2d(&tmWqkv, a.wqkv, (uint64_t)3 * a.dim, (uint64_t)a.dim, (uint64_t)a.dim * 2, P.g_qkv.BN >> (pair ? 1 : 0), 64, true))) return rc;

  // the instrumented variant only when one of its features is asked for (see EXTRA above)
  const bool extra = dbg != 0 || fast_silu == 0 || P.pf_budget > 0 || P.pf_flat != nullptr;
  const bool narrow = P.g_wo.BN <= 64 && P.g_w13.BN <= 64 && P.g_w2.BN <= 64 && P.g_qkv.BN <= 64;   // ring layout by the widest column tile in use
#define VLG_LT_LAUNCH(ST_, SB_, PAIR_, EXTRA_, CLUSTER_)                                                                          \
  {                                                                                                                               \
    auto kern = layer_tail_kernel<ST_, SB_, PAIR_, EXTRA_>;                                                                       \
    const int smem = lt_smem_bytes(ST_, SB_);                                                                                     \
    static PerDeviceOnce attr_once;                                                                                               \
    VLG_CUDA(attr_once([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); }));           \
    if (CLUSTER_ > 1) VLG_CUDA(launch_kernel_cluster(kern, dim3(sms), dim3(kLtThreads), (size_t)smem, st, CLUSTER_, tmAttn, tmXn, tmFfn, tmWo, tmW13, tmW2, tmWqkv, P)); \
    else VLG_CUDA(launch_kernel(kern, dim3(sms), dim3(kLtThreads), (size_t)smem, st, tmAttn, tmXn, tmFfn, tmWo, tmW13, tmW2, tmWqkv, P)); \
  }
  if (pair) {
    if (extra) VLG_LT_LAUNCH(kLtStagesPair, kLtStageBytesPair, true, true, 2)
    else VLG_LT_LAUNCH(kLtStagesPair, kLtStageBytesPair, true, false, 2)
  } else if (narrow) {
    if (extra) VLG_LT_LAUNCH(kLtStagesNarrow, kLtStageBytesNarrow, false, true, 1)
    else VLG_LT_LAUNCH(kLtStagesNarrow, kLtStageBytesNarrow, false, false, 1)
  } else {
    if (extra) VLG_LT_LAUNCH(kLtStages, kLtStageBytesWide, false, true, 1)
    else VLG_LT_LAUNCH(kLtStages, kLtStageBytesWide, false, false, 1)
  }
#undef VLG_LT_LAUNCH
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg

// instrumentation: phase timestamps (ns), [160 CTAs][32 slots], of the last full layer-tail launch run with VLG_LT_DEBUG=1
extern "C" int vlg_debug_lt_timestamps(unsigned long long* out_host) {
  VLG_CUDA(cudaMemcpyFromSymbol(out_host, vlg::g_lt_ts, sizeof(unsigned long long) * 160 * 32));
  return VLG_OK;
}
