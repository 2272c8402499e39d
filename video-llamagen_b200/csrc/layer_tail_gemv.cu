// The layer tail (everything between two attentions of the bf16 decode step, see layer_tail.cu) for 1-2 rows: a single
// image with guidance, the t2v loop at batch 1 (BASELINE cfg-4, B = 1).
//
// At that size the tensor-core kernel is pure latency -- its six phases, five grid barriers and split-K exchanges take 25 us
// whatever the row count, for 6 us worth of weights -- and a matrix-vector product needs no tensor core.  Here a warp owns PAIRS
// of output columns (a RoPE pair, a (w1 x, w3 x) pair); every CTA owns a contiguous block of pairs of each matrix, and a
// producer thread streams those weight rows through a 12-stage, 192 KB shared-memory ring with bulk copies
// (cp.async.bulk, one per stage of up to 16 KB) that runs ahead ACROSS the phases -- weights depend on nothing, so while the
// grid sits in a barrier the next phase's rows are already landing (a first version loaded each pair's rows from global
// memory into registers one pair ahead: every pair cost a dependent HBM round trip, 714 against 700 latents/s at cfg-4 B = 1).
// Consumer warp w takes every 8th stage and multiplies it against activation rows it keeps packed in registers; every GEMV
// owns all of K, so there are no partials and no row phases: the RMSNorm in front of w1|w3 and wqkv is recomputed by every
// warp (2 rows x dim values) on the way in.  Three grid barriers instead of five:
//
//   wo + residual | B | RMSNorm -> w1|w3 -> SwiGLU | B | w2 + residual | B | RMSNorm -> wqkv -> RoPE + q / KV-cache write
//
// Rounding points are those of the per-op kernels (epilogue.cuh / elementwise.cu): GEMV results round to bf16
// before the elementwise tail, RMSNorm rounds before and after the weight multiply, RoPE products round separately.
#include "lt_common.cuh"

namespace vlg {

constexpr int kGvWarps = 8;                          // consumer warps; warp kGvWarps is the bulk-copy producer
constexpr int kGvThreads = (kGvWarps + 1) * 32;
constexpr int kGvStageBytes = 16384;                 // one stage: as many whole column pairs (2 rows of K bf16) as fit, at least one
constexpr int kGvStages = 12;
constexpr int kGvActBytes = 16384;                   // the SwiGLU output of all rows, staged once per CTA for the w2 phase (2 x hidden <= 4096 bf16)
constexpr int kGvMaxRows = 2;


struct GvParams {
  int rows, dim, hidden, has_front, has_qkv;
  int dbg;                   // VLG_LT_DEBUG=1: globaltimer stamps of every CTA's warp 0 (tools/lt_phases.py)
  const __nv_bfloat16 *attn, *wo, *w13, *w2, *wqkv, *norm1_w, *norm2_w;
  __nv_bfloat16 *h, *xn, *ffn;
  float eps;
  EpiParams rope;
  unsigned* bar;             // [0] barrier arrivals, [1] exit arrivals (the layer-tail counter region)
  volatile unsigned* err;
};

// phase timestamps (ns) of the last full launch, [160 CTAs][32 slots], read back through vlg_debug_lt_timestamps
__device__ unsigned long long g_gv_ts[160 * 32];
#define GV_TS(i)                                                                                      \
  do {                                                                                                \
    if (P.dbg && P.has_front && P.has_qkv && threadIdx.x == 0 && blockIdx.x < 160) {                  \
      unsigned long long t_;                                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                                          \
      g_gv_ts[blockIdx.x * 32 + (i)] = t_;                                                            \
    }                                                                                                 \
  } while (0)

__device__ __forceinline__ void gv_grid_barrier(unsigned* ctr, unsigned& target, volatile unsigned* err) {
  target += gridDim.x;
  asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");      // the consumer warps (the producer never stops for the grid)
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
  asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
}

__device__ __forceinline__ void gv_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 gv_lds128(const void* p) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
  return r;
}
struct GvSmem { uint64_t full[kGvStages], empty[kGvStages]; };
// this CTA's contiguous block of the `pairs` column pairs of a matrix (balanced to within one pair)
__device__ __forceinline__ void gv_range(int pairs, int& p0, int& p1) {
  const int base = pairs / (int)gridDim.x, rem = pairs % (int)gridDim.x, c = blockIdx.x;
  p0 = c * base + min(c, rem);
  p1 = p0 + base + (c < rem ? 1 : 0);
}
__device__ __forceinline__ int gv_pairs_per_stage(int K) { const int n = kGvStageBytes / (4 * K); return n < 1 ? 1 : n; }

__device__ __forceinline__ void gv_unpack(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}

// The activation rows of a phase, ONCE per CTA into shared memory (all consumer threads, one L2 round trip).  When every warp
// of the grid fetched the rows itself, 1184 warps asked the same 40 cache lines for the same bytes at the same moment.
// Row r lies at acts + r * 2 K bytes; with `norm_w` the rows are RMS-normalised in place by warps 0 .. rows-1 before anybody
// reads them: T(T(h * rsqrt(mean(h^2) + eps)) * w)   (gpt.py:143-148).
// Loops over K are ROLLED on purpose, here and in gv_phase: every phase runs once per launch, so its instructions are fetched
// cold -- the fully unrolled first version of this kernel was 155 KB of SASS and took 6 us longer with 2 rows than with 1.
template <int R>
__device__ __forceinline__ void gv_stage_rows(const __nv_bfloat16* src, int K, int rows, const __nv_bfloat16* norm_w, float eps, uint8_t* acts,
                                              int warp, int lane) {
  const int per_row = K >> 3;
  for (int i = threadIdx.x; i < R * per_row; i += kGvWarps * 32) {
    const int r = i / per_row, c = i - r * per_row;
    *reinterpret_cast<uint4*>(acts + (size_t)i * 16) = ldg128_cg(src + (size_t)min(r, rows - 1) * K + c * 8);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
  if (norm_w == nullptr) return;
  if (warp < R) {
    uint8_t* row = acts + (size_t)warp * 2 * K;
    float ss = 0.f;
#pragma unroll 1
    for (int c = lane; c < per_row; c += 32) {
      float f[8];
      gv_unpack(gv_lds128(row + c * 16), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = __fadd_rn(ss, __fmul_rn(f[i], f[i]));
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(__fdiv_rn(ss, (float)K) + eps);
#pragma unroll 1
    for (int c = lane; c < per_row; c += 32) {
      float f[8], g[8];
      gv_unpack(gv_lds128(row + c * 16), f);
      gv_unpack(__ldg(reinterpret_cast<const uint4*>(norm_w + c * 8)), g);
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float y0 = __fmul_rn(f[2 * i], rs), y1 = __fmul_rn(f[2 * i + 1], rs);
        rnd2_bf16(y0, y1);
        o[i] = pack_bf16(__fmul_rn(y0, g[2 * i]), __fmul_rn(y1, g[2 * i + 1]));
      }
      *reinterpret_cast<uint4*>(row + c * 16) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
}

template <int R>
__device__ __forceinline__ void gv_reduce(float (&acc)[R][2]) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    acc[r][0] = warp_sum(acc[r][0]);
    acc[r][1] = warp_sum(acc[r][1]);
  }
}

enum { GV_RESID = 0, GV_SWIGLU = 1, GV_ROPE = 2 };

// what lane r needs from global memory to finalise row r of the pair (n, n + 1): requested BEFORE the dot product, so the
// round trip overlaps it (a RoPE finish is two dependent loads -- position, then table entry -- a residual finish one)
struct GvPre { uint32_t hv; float2 cs; };
template <int MODE>
__device__ __forceinline__ GvPre gv_prefetch(const GvParams& P, int n, int pos, int lane) {
  GvPre pre{0u, make_float2(1.f, 0.f)};
  if (lane >= P.rows) return pre;
  if constexpr (MODE == GV_RESID) {
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(pre.hv) : "l"(P.h + (size_t)lane * P.dim + n));
  } else if constexpr (MODE == GV_ROPE) {
    const EpiParams& E = P.rope;
    const int sec = n / E.dim, cc = n - sec * E.dim, dd = cc % E.D;
    if (sec < 2) pre.cs = __ldg(reinterpret_cast<const float2*>(E.rope) + (size_t)pos * (E.D >> 1) + (dd >> 1));
  }
  return pre;
}

// lane r finalises row r of the pair (n, n + 1)
template <int MODE, int R>
__device__ __forceinline__ void gv_finish(const GvParams& P, int n, int pos, const GvPre& pre, const float (&acc)[R][2], int lane) {
  float a0 = acc[0][0], a1 = acc[0][1];
#pragma unroll
  for (int r = 1; r < R; ++r) {
    a0 = lane == r ? acc[r][0] : a0;
    a1 = lane == r ? acc[r][1] : a1;
  }
  if (lane >= P.rows) return;
  const int r = lane;
  rnd2_bf16(a0, a1);                                   // the nn.Linear result as a bf16 tensor
  if constexpr (MODE == GV_RESID) {                    // gpt.py:257-258
    const uint32_t o = pack_bf16(__fadd_rn(bf16lo(pre.hv), a0), __fadd_rn(bf16hi(pre.hv), a1));
    asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(P.h + (size_t)r * P.dim + n), "r"(o) : "memory");
  } else if constexpr (MODE == GV_SWIGLU) {            // gpt.py:167, w1 / w3 rows interleaved
    const float s = rnd<__nv_bfloat16>(silu_f(a0));
    P.ffn[(size_t)r * P.hidden + (n >> 1)] = __float2bfloat16_rn(__fmul_rn(s, a1));
  } else {                                             // gpt.py:215-227, 423-433, 177-185
    const EpiParams& E = P.rope;
    const int sec = n / E.dim, cc = n - sec * E.dim, head = cc / E.D, dd = cc - head * E.D;
    float o0 = a0, o1 = a1;
    if (sec < 2) {
      o0 = __fsub_rn(__fmul_rn(a0, pre.cs.x), __fmul_rn(a1, pre.cs.y));
      o1 = __fadd_rn(__fmul_rn(a1, pre.cs.x), __fmul_rn(a0, pre.cs.y));
    }
    __nv_bfloat16* dst;
    if (sec == 0) dst = static_cast<__nv_bfloat16*>(E.q_out) + (size_t)r * E.dim + cc;
    else dst = static_cast<__nv_bfloat16*>(sec == 1 ? E.k_cache : E.v_cache) + (((size_t)r * E.H + head) * E.S_max + pos) * E.D + dd;
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16(o0, o1);
  }
}

// One GEMV phase of a consumer warp: the stages of this CTA's block of pairs arrive in order; stage number gs (counted over
// the whole launch, in step with the producer) belongs to warp gs % kGvWarps.  Up to kGP pairs of a stage are multiplied
// together (4 * kGP independent accumulation chains, the activation chunk unpacked once for all of them) against the rows
// in shared memory.
constexpr int kGP = 3;
template <int MODE, int R>
__device__ __forceinline__ void gv_phase(const GvParams& P, int N, int K, const uint8_t* acts, int pos, uint32_t& gs, uint8_t* ring, GvSmem& sm,
                                         int warp, int lane) {
  int p0, p1;
  gv_range(N >> 1, p0, p1);
  const int pps = gv_pairs_per_stage(K), per_row = K >> 3;
  for (int p = p0; p < p1; p += pps, ++gs) {
    if ((int)(gs % kGvWarps) != warp) continue;
    const uint32_t st = gs % kGvStages, it = gs / kGvStages;
    const int np = min(pps, p1 - p);
    bool waited = false;
#pragma unroll 1
    for (int q0 = 0; q0 < np; q0 += kGP) {
      const int g = min(kGP, np - q0);
      GvPre pre[kGP];
#pragma unroll
      for (int j = 0; j < kGP; ++j)
        if (j < g) pre[j] = gv_prefetch<MODE>(P, 2 * (p + q0 + j), pos, lane);
      if (!waited) { mbar_wait(&sm.full[st], it & 1); waited = true; }
      const uint8_t* wbase = ring + st * kGvStageBytes + (size_t)q0 * 4 * K;
      float acc[kGP][R][2];
#pragma unroll
      for (int j = 0; j < kGP; ++j)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[j][r][0] = acc[j][r][1] = 0.f;
#pragma unroll 1
      for (int c = lane; c < per_row; c += 32) {
        float f[R][8];
#pragma unroll
        for (int r = 0; r < R; ++r) gv_unpack(gv_lds128(acts + (size_t)r * 2 * K + c * 16), f[r]);
#pragma unroll
        for (int j = 0; j < kGP; ++j) {
          if (j >= g) continue;
          float w0[8], w1[8];
          gv_unpack(gv_lds128(wbase + (size_t)j * 4 * K + c * 16), w0);
          gv_unpack(gv_lds128(wbase + (size_t)j * 4 * K + 2 * K + c * 16), w1);
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              acc[j][r][0] = fmaf(f[r][i], w0[i], acc[j][r][0]);
              acc[j][r][1] = fmaf(f[r][i], w1[i], acc[j][r][1]);
            }
        }
      }
#pragma unroll
      for (int j = 0; j < kGP; ++j)
        if (j < g) gv_reduce<R>(acc[j]);
#pragma unroll
      for (int j = 0; j < kGP; ++j)
        if (j < g) gv_finish<MODE, R>(P, 2 * (p + q0 + j), pos, pre[j], acc[j], lane);
    }
    __syncwarp();                                     // every lane is done reading the stage
    if (lane == 0) mbar_arrive(&sm.empty[st]);
  }
}

// producer: this CTA's block of pairs of one matrix, one bulk copy per stage (the rows of consecutive pairs are contiguous)
__device__ __forceinline__ void gv_stream(const __nv_bfloat16* W, int N, int K, uint32_t& gs, uint8_t* ring, GvSmem& sm) {
  int p0, p1;
  gv_range(N >> 1, p0, p1);
  const int pps = gv_pairs_per_stage(K);
  for (int p = p0; p < p1; p += pps, ++gs) {
    const uint32_t st = gs % kGvStages, it = gs / kGvStages;
    const uint32_t bytes = (uint32_t)min(pps, p1 - p) * 4u * (uint32_t)K;
    mbar_wait(&sm.empty[st], (it & 1) ^ 1);
    mbar_expect_tx(&sm.full[st], bytes);
    gv_bulk_g2s(ring + st * kGvStageBytes, W + (size_t)2 * p * K, bytes, &sm.full[st]);
  }
}

template <int R>
__global__ void __launch_bounds__(kGvThreads, 1) layer_tail_gemv_kernel(const __grid_constant__ GvParams P) {
  extern __shared__ uint8_t gv_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(gv_smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* acts = ring + kGvStages * kGvStageBytes;
  GvSmem& sm = *reinterpret_cast<GvSmem*>(acts + kGvActBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.dim, hid = P.hidden;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kGvStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  uint32_t gs = 0;
  if (warp == kGvWarps) {
    // ---- producer: the weights depend on nothing -- neither on the previous kernel nor on this kernel's barriers
    if (elect_one()) {
      if (P.has_front) {
        gv_stream(P.wo, d, d, gs, ring, sm);
        gv_stream(P.w13, 2 * hid, d, gs, ring, sm);
        gv_stream(P.w2, d, hid, gs, ring, sm);
      }
      if (P.has_qkv) gv_stream(P.wqkv, 3 * d, d, gs, ring, sm);
    }
    return;
  }
  // ---- consumers
  unsigned bar_target = 0;
  GV_TS(0);
  pdl_wait();
  GV_TS(1);
  if (P.has_front) {
    gv_stage_rows<R>(P.attn, d, P.rows, nullptr, 0.f, acts, warp, lane);
    gv_phase<GV_RESID, R>(P, d, d, acts, 0, gs, ring, sm, warp, lane);
    GV_TS(2);
    gv_grid_barrier(P.bar, bar_target, P.err);
    GV_TS(3);
    gv_stage_rows<R>(P.h, d, P.rows, P.norm1_w, P.eps, acts, warp, lane);
    GV_TS(4);
    gv_phase<GV_SWIGLU, R>(P, 2 * hid, d, acts, 0, gs, ring, sm, warp, lane);
    GV_TS(5);
    gv_grid_barrier(P.bar, bar_target, P.err);
    GV_TS(6);
    gv_stage_rows<R>(P.ffn, hid, P.rows, nullptr, 0.f, acts, warp, lane);
    GV_TS(7);
    gv_phase<GV_RESID, R>(P, d, hid, acts, 0, gs, ring, sm, warp, lane);
    GV_TS(8);
    gv_grid_barrier(P.bar, bar_target, P.err);
    GV_TS(9);
  }
  const int pos = P.has_qkv ? (P.rope.pos_ptr ? *P.rope.pos_ptr : P.rope.pos_imm) : 0;
  gv_stage_rows<R>(P.h, d, P.rows, P.norm2_w, P.eps, acts, warp, lane);      // (behind a grid barrier, or first thing in the launch)
  GV_TS(10);
  if (P.has_qkv) {
    gv_phase<GV_ROPE, R>(P, 3 * d, d, acts, pos, gs, ring, sm, warp, lane);
    GV_TS(11);
  } else if (blockIdx.x == 0) {                          // after the last layer: the head GEMM reads xn from global memory
    for (int i = threadIdx.x; i < P.rows * (d >> 3); i += kGvWarps * 32)
      *reinterpret_cast<uint4*>(P.xn + (size_t)i * 8) = *reinterpret_cast<const uint4*>(acts + (size_t)i * 16);
  }

  // leave the counters at zero for the next launch
  asm volatile("bar.sync 1, %0;" ::"n"(kGvWarps * 32) : "memory");
  if (threadIdx.x == 0) {
    const unsigned old = atomicAdd(P.bar + 1, 1u);
    if (old == gridDim.x - 1) {
      P.bar[0] = 0;
      P.bar[1] = 0;
      __threadfence();
    }
  }
}

bool layer_tail_gemv_supported(int rows, int dim, int hidden) {
  static const bool off = [] { const char* e = getenv("VLG_LT_GEMV"); return e && e[0] == '0'; }();
  return !off && rows >= 1 && rows <= kGvMaxRows && dim % 256 == 0 && hidden % 256 == 0 && 4 * hidden <= kGvStageBytes && kGvMaxRows * 2 * hidden <= kGvActBytes &&
         kGvMaxRows * 2 * dim <= kGvActBytes;
}

int launch_layer_tail_gemv(const LtLaunch& a, cudaStream_t st) {
  VLG_REQUIRE(layer_tail_gemv_supported(a.rows, a.dim, a.hidden), "layer_tail_gemv: unsupported shape rows=%d dim=%d hidden=%d", a.rows, a.dim, a.hidden);
  VLG_REQUIRE(a.wo || a.wqkv, "layer_tail_gemv: nothing to do");
  VLG_REQUIRE(a.rope.D % 2 == 0 && a.dim % 2 == 0, "layer_tail_gemv: odd head_dim");
  GvParams P{};
  P.rows = a.rows; P.dim = a.dim; P.hidden = a.hidden;
  P.has_front = a.wo != nullptr;
  P.has_qkv = a.wqkv != nullptr;
  P.attn = static_cast<const __nv_bfloat16*>(a.attn);
  P.wo = static_cast<const __nv_bfloat16*>(a.wo); P.w13 = static_cast<const __nv_bfloat16*>(a.w13);
  P.w2 = static_cast<const __nv_bfloat16*>(a.w2); P.wqkv = static_cast<const __nv_bfloat16*>(a.wqkv);
  P.norm1_w = static_cast<const __nv_bfloat16*>(a.norm1_w); P.norm2_w = static_cast<const __nv_bfloat16*>(a.norm2_w);
  P.h = static_cast<__nv_bfloat16*>(a.h); P.xn = static_cast<__nv_bfloat16*>(a.xn); P.ffn = static_cast<__nv_bfloat16*>(a.ffn);
  P.eps = a.eps;
  P.rope = a.rope;
  P.bar = a.bar;
  P.err = a.err;
  static const int dbg = [] { const char* e = getenv("VLG_LT_DEBUG"); return e ? atoi(e) : 0; }();
  P.dbg = dbg;
  const dim3 grid(num_sms()), block(kGvThreads);
  const size_t smem = (size_t)kGvStages * kGvStageBytes + kGvActBytes + sizeof(GvSmem) + 128;
  static PerDeviceOnce once1, once2;
  if (a.rows == 1) {
    auto kern = layer_tail_gemv_kernel<1>;
    VLG_CUDA(once1([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }));
    VLG_CUDA(launch_kernel(kern, grid, block, smem, st, P));
  } else {
    auto kern = layer_tail_gemv_kernel<2>;
    VLG_CUDA(once2([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }));
    VLG_CUDA(launch_kernel(kern, grid, block, smem, st, P));
  }
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

// instrumentation: the stamps of the last full GEMV-tail launch (vlg_debug_lt_timestamps adds them behind the tensor-core
// kernel's: out[160 * 32 ...])
int layer_tail_gemv_timestamps(unsigned long long* out_host) {
  VLG_CUDA(cudaMemcpyFromSymbol(out_host, g_gv_ts, sizeof(unsigned long long) * 160 * 32));
  return VLG_OK;
}

}  // namespace vlg
