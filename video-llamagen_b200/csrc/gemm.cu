// out[rows, N] = x[rows, K] . w[N, K]^T (nn.Linear without bias) with fused epilogues (epilogue.cuh).
//
//  * bf16: tcgen05 tensor cores.  One CTA owns a [128 x BN] output tile: rows ride the 128 TMEM lanes
//    (TMA zero-fills rows >= `rows`), BN weight rows are the N operand.  K is streamed through an
//    8-stage TMA ring (A = activations 128x64, B = weights BNx64, both K-major, 128B swizzle); one
//    elected thread issues tcgen05.mma (M=128, N=BN, K=16) with the fp32 accumulator in TMEM; four
//    epilogue warps read it back with tcgen05.ld (lane == output row) and apply the epilogue.
//    The decode GEMMs are weight-streaming bound (M <= 256): BN is small so that N/BN CTAs spread the
//    weight stream over the SMs; activations are re-read from L2 (evict-last), weights evict-first.
//  * fp32: CUDA-core FMA only (no TF32), fixed summation order -- the exact path behind the greedy
//    bit-exact contract.  One warp owns 8 output columns, lanes split K, 8 rows per pass.
#include <stdlib.h>

#include "common.cuh"
#include "epilogue.cuh"

namespace vlg {

// ============================================================================ tcgen05 bf16 GEMM
constexpr int kMmaWarps = 2;       // MMA issuers: k-blocks alternate between them, one TMEM accumulator each
constexpr int kGemmThreads = 32 * (1 + kMmaWarps + 4);  // warp 0: TMA producer, warps 1..kMmaWarps: MMA issuers (warp 1 owns TMEM), 4 epilogue warps
constexpr int kBK = 64;            // 64 bf16 = one 128-byte swizzle row

// BM = 128: output rows ride all 128 TMEM lanes.  BM = 64: "half subpartition" accumulator layout -- row m lives in
// lane (m % 16) + 32 * (m / 16) (cute::UMMA::tmem_frg, M_MMA == 64) -- used where the activation tile, not the weight
// stream, bounds the CTA (each CTA ingests the whole [BM x K] activation slab through its 64 B/clk L2 port).
template <int BM, int BN> struct GemmCfg {
  static constexpr int kABytes = BM * kBK * 2;
  static constexpr int kBBytes = (BN * kBK * 2 + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // The k-loop is latency-bound (one TMA round trip is ~1.3 us against ~0.05 us of MMA per k-block): the ring is as deep
  // as shared memory allows so that (nearly) all of K is in flight at once.
  static constexpr int kStagesRaw = (200 * 1024 / kStageBytes) < 24 ? (200 * 1024 / kStageBytes) : 24;
  static constexpr int kStages = kStagesRaw / kMmaWarps * kMmaWarps;   // a stage is always consumed by the same MMA warp
  static constexpr int kAccs = kMmaWarps;   // one accumulator per MMA-issuing warp, summed in the epilogue
  static constexpr int kTmemCols = kAccs * BN < 32 ? 32 : kAccs * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 512 /*barriers*/;
};

// phase timestamps of CTA 0 (globaltimer ns), written when dbg & 4: microbenchmark instrumentation only
__device__ unsigned long long g_gemm_ts[16];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GEMM_TS(i) do { if ((dbg & 4) && blockIdx.x == 0 && blockIdx.y == 0) g_gemm_ts[i] = gtime(); } while (0)

constexpr int kNormCtas = 64;       // CTAs that share the fused RMSNorm prologue (always the first ones dispatched)

// RMSNorm of row r by the 128 epilogue threads (te = 0..127); bit-identical to rmsnorm_kernel<bf16> (same partial sums).
__device__ __forceinline__ void fused_rmsnorm_row(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ w,
                                                  __nv_bfloat16* __restrict__ dst, int K, float eps, int te, float* s_part) {
  const int nvec = K / 8;
  float v[4][8], g[4][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = te + i * 128;
    if (vi < nvec) {
      load8<__nv_bfloat16>(w + vi * 8, g[i]);
      load8<__nv_bfloat16>(src + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = __fadd_rn(ss, __fmul_rn(v[i][j], v[i][j]));
    }
  }
  ss = warp_sum(ss);
  asm volatile("bar.sync 2, 128;" ::: "memory");          // previous row's readers of s_part are done
  if ((te & 31) == 0) s_part[te >> 5] = ss;
  asm volatile("bar.sync 2, 128;" ::: "memory");
  const float tot = ((s_part[0] + s_part[1]) + s_part[2]) + s_part[3];
  const float rs = rsqrtf(__fdiv_rn(tot, (float)K) + eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = te + i * 128;
    if (vi < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __fmul_rn(rnd<__nv_bfloat16>(__fmul_rn(v[i][j], rs)), g[i][j]);
      store8<__nv_bfloat16>(dst + vi * 8, o);
    }
  }
}

template <int BM, int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tc_kernel(const __grid_constant__ TensorMap tmA, const __grid_constant__ TensorMap tmB, int rows, int K,
                    EpiParams P, int dbg, uint32_t a_bytes) {
  using Cfg = GemmCfg<BM, BN>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_norm_part[4];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* acc_ready = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int nkb = K / kBK;
  if (threadIdx.x == 0) GEMM_TS(0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_ready, kMmaWarps);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;
  if (threadIdx.x == 0) GEMM_TS(1);

  pdl_launch_dependents();
  if (warp == 0) {
    if (elect_one()) {   // elect.sync: the compiler then knows one thread is active and feeds the uniform datapath directly
      // The weight tiles do not depend on the previous kernel: start streaming them for the first ring pass before
      // waiting for its output, so the weight stream overlaps the predecessor's tail.
      const int pre = nkb < Cfg::kStages ? nkb : Cfg::kStages;
      // dbg (microbenchmarks only): bit 0 skips the activation loads, bit 1 the weight loads
      const uint32_t tx = ((dbg & 1) ? 0 : a_bytes) + ((dbg & 2) ? 0 : BN * kBK * 2);
      for (int kb = 0; kb < pre; ++kb) {
        if (tx) mbar_expect_tx(&full[kb], tx); else mbar_arrive(&full[kb]);
        if (!(dbg & 2)) tma_load_2d_hint(smem + kb * Cfg::kStageBytes + Cfg::kABytes, &tmB, &full[kb], kb * kBK, n0, kEvictFirst);
      }
      GEMM_TS(2);
      pdl_wait();
      if (P.norm_flag != nullptr) {           // fused RMSNorm: A is being written by the first CTAs of this very grid
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned need = total < (unsigned)kNormCtas ? total : (unsigned)kNormCtas;
        uint32_t spins = 0;
        unsigned seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(P.norm_flag) : "memory");
          if (++spins > (1u << 24)) { printf("vlg: fused rmsnorm flag timeout (block %d)\n", blockIdx.x); __trap(); }
        } while (seen < need);
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of other SMs -> this SM's TMA reads
      }
      GEMM_TS(3);
      for (int kb = 0; kb < pre; ++kb)
        if (!(dbg & 1)) tma_load_2d_hint(smem + kb * Cfg::kStageBytes, &tmA, &full[kb], kb * kBK, m0, kEvictLast);
      for (int kb = pre; kb < nkb; ++kb) {
        const int s = kb % Cfg::kStages, it = kb / Cfg::kStages;
        mbar_wait(&empty[s], (it & 1) ^ 1);
        uint8_t* a = smem + s * Cfg::kStageBytes;
        if (tx) mbar_expect_tx(&full[s], tx); else mbar_arrive(&full[s]);
        if (!(dbg & 1)) tma_load_2d_hint(a, &tmA, &full[s], kb * kBK, m0, kEvictLast);
        if (!(dbg & 2)) tma_load_2d_hint(a + Cfg::kABytes, &tmB, &full[s], kb * kBK, n0, kEvictFirst);
      }
    }
  } else if (warp <= kMmaWarps) {
    // Two issuing threads: a single thread sustains only ~1 tcgen05.mma per 75 cycles (measured, independent of the tile
    // shape), well below what the tensor pipe accepts for these narrow tiles.  dbg & 16 routes everything through warp 1.
    const int j = warp - 1;
    const bool solo = (dbg & 16) != 0;
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      const uint32_t acc = tmem_acc + (solo ? 0 : j * BN);
      bool first = true;
      if (!(solo && j > 0)) {
        for (int kb = solo ? 0 : j; kb < nkb; kb += solo ? 1 : kMmaWarps) {
          const int s = kb % Cfg::kStages, it = kb / Cfg::kStages;
          mbar_wait(&full[s], it & 1);
          if (kb == 0) GEMM_TS(4);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)  // +32 bytes along K inside the swizzle row = +2 in the address field
            umma_bf16(acc, adesc + 2 * k, bdesc + 2 * k, idesc, !(first && k == 0));
          first = false;
          umma_commit(&empty[s]);             // frees the smem stage once these MMAs retire
        }
      }
      umma_commit(acc_ready);
      if (j == 0) GEMM_TS(5);
    }
  } else {
    const int quad = warp & 3;              // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
    const int r = BM == 128 ? m0 + quad * 32 + lane : (lane < 16 ? m0 + quad * 16 + lane : 0x7fffffff);
    const int n_acc = (dbg & 16) ? 1 : (nkb < kMmaWarps ? nkb : kMmaWarps);
    pdl_wait();                             // the epilogue reads predecessor outputs (residual, position)
    if (P.norm_flag != nullptr) {
      const unsigned total = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
      const unsigned n_norm = total < (unsigned)kNormCtas ? total : (unsigned)kNormCtas;
      if (bid < n_norm) {
        const int te = threadIdx.x - 32 * (1 + kMmaWarps);
        for (int rr = bid; rr < rows; rr += n_norm)
          fused_rmsnorm_row(static_cast<const __nv_bfloat16*>(P.norm_src) + (size_t)rr * K,
                            static_cast<const __nv_bfloat16*>(P.norm_weight),
                            static_cast<__nv_bfloat16*>(P.norm_dst) + (size_t)rr * K, K, P.norm_eps, te, s_norm_part);
        __threadfence();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (te == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(P.norm_flag) : "memory");
      }
    }
    mbar_wait(acc_ready, 0);
    if (threadIdx.x == 32 * (1 + kMmaWarps)) GEMM_TS(6);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < BN; c += 16) {
      float v[16];
      tmem_ld16(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
#pragma unroll
      for (int a = 1; a < Cfg::kAccs; ++a) {
        if (a < n_acc) {                    // warp-uniform
          float u[16];
          tmem_ld16(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + a * BN + c, u);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += u[i];
        }
      }
      if (r < rows) {
        epi8<EPI, __nv_bfloat16>(r, n0 + c, v, P);
        epi8<EPI, __nv_bfloat16>(r, n0 + c + 8, v + 8, P);
      }
    }
  }
  if (threadIdx.x == 32 * (1 + kMmaWarps)) GEMM_TS(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_acc);
  if (threadIdx.x == 32) GEMM_TS(8);
}

template <int BM, int BN, int EPI>
static int launch_tc(const void* x, const void* w, int rows, int N, int K, const EpiParams& P, cudaStream_t st) {
  using Cfg = GemmCfg<BM, BN>;
  TensorMap tmA, tmB;
  // the activation box is as tall as the tensor when that is less than a tile: TMA would zero-fill the rest at the cost of
  // real bytes (the MMA then reads stale shared memory for those rows; their accumulator lanes are never read back)
  const uint32_t a_rows = rows < BM ? (uint32_t)((rows + 7) / 8 * 8) : (uint32_t)BM;
  int rc = make_tmap_bf16_2d(&tmA, x, (uint64_t)rows, (uint64_t)K, (uint64_t)K * 2, a_rows, kBK, true);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, w, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, BN, kBK, true);
  if (rc) return rc;
  auto kern = gemm_bf16_tc_kernel<BM, BN, EPI>;
  static PerDeviceOnce attr_once;
  VLG_CUDA(attr_once([&] { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes); }));
  dim3 grid(N / BN, (rows + BM - 1) / BM);
  static const int dbg = [] { const char* e = getenv("VLG_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  VLG_CUDA(launch_kernel(kern, grid, dim3(kGemmThreads), Cfg::kSmemBytes, st, tmA, tmB, rows, K, P, dbg, a_rows * (uint32_t)(kBK * 2)));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

// ============================================================================ exact fp32 GEMM
constexpr int kF32Warps = 4;

template <int EPI>
__global__ void __launch_bounds__(kF32Warps * 32)
gemm_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, int rows, int N, int K, EpiParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * kF32Warps + warp) * 8;
  pdl_launch_dependents();
  pdl_wait();
  if (n0 >= N) return;
  for (int r0 = 0; r0 < rows; r0 += 8) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      float4 wv[8], xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = *reinterpret_cast<const float4*>(w + (size_t)(n0 + j) * K + k);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int r = min(r0 + i, rows - 1);
        xv[i] = *reinterpret_cast<const float4*>(x + (size_t)r * K + k);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = acc[i][j];
          a = __fmaf_rn(xv[i].x, wv[j].x, a);
          a = __fmaf_rn(xv[i].y, wv[j].y, a);
          a = __fmaf_rn(xv[i].z, wv[j].z, a);
          a = __fmaf_rn(xv[i].w, wv[j].w, a);
          acc[i][j] = a;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = warp_sum(acc[i][j]);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (lane == i && r0 + i < rows) epi8<EPI, float>(r0 + i, n0, acc[i], P);
  }
}


// rows <= 8 (every position of the fp32 decode loop): the same sums in the same order -- the results are bit-identical to
// gemm_f32_kernel's -- spread over four times as many warps.  The fp32 path streams every weight once per position, so what
// matters is bytes in flight; with 8 columns per warp a [2048 x 2048] layer leaves 256 warps for the whole chip (measured:
// 1.6 TB/s over a GPT-1B trunk position).  Here four warps share an 8-column group, two columns each, the group's sums meet
// in shared memory and one warp runs the 8-column epilogue.
constexpr int kF32FewWarps = 8;      // two 8-column groups per CTA

template <int EPI>
__global__ void __launch_bounds__(kF32FewWarps * 32)
gemm_f32_few_kernel(const float* __restrict__ x, const float* __restrict__ w, int rows, int N, int K, EpiParams P) {
  __shared__ float sums[2][8][8];      // [group][row][column]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp >> 2, sub = warp & 3;
  const int n0 = (blockIdx.x * 2 + grp) * 8, nc = n0 + sub * 2;
  pdl_launch_dependents();
  pdl_wait();
  const bool live = n0 < N;
  if (live) {
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.f;
    const float* w0 = w + (size_t)nc * K;
    const float* w1 = w0 + K;
#pragma unroll 4
    for (int k = lane * 4; k < K; k += 128) {
      const float4 a = *reinterpret_cast<const float4*>(w0 + k), b = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i >= rows) continue;
        const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)i * K + k);
        float s = acc[i][0], t = acc[i][1];
        s = __fmaf_rn(xv.x, a.x, s); s = __fmaf_rn(xv.y, a.y, s); s = __fmaf_rn(xv.z, a.z, s); s = __fmaf_rn(xv.w, a.w, s);
        t = __fmaf_rn(xv.x, b.x, t); t = __fmaf_rn(xv.y, b.y, t); t = __fmaf_rn(xv.z, b.z, t); t = __fmaf_rn(xv.w, b.w, t);
        acc[i][0] = s; acc[i][1] = t;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i >= rows) continue;
      const float s = warp_sum(acc[i][0]), t = warp_sum(acc[i][1]);
      if (lane == 0) { sums[grp][i][sub * 2] = s; sums[grp][i][sub * 2 + 1] = t; }
    }
  }
  __syncthreads();
  if (live && sub == 0 && lane < rows) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = sums[grp][lane][j];
    epi8<EPI, float>(lane, n0, v, P);
  }
}

template <int EPI>
static int launch_f32(const void* x, const void* w, int rows, int N, int K, const EpiParams& P, cudaStream_t st) {
  int groups = N / 8;
  static const bool few_off = [] { const char* e = getenv("VLG_F32_FEW"); return e && e[0] == '0'; }();    // A/B
  if (rows <= 8 && !few_off) {
    VLG_CUDA(launch_kernel(gemm_f32_few_kernel<EPI>, dim3((groups + 1) / 2), dim3(kF32FewWarps * 32), 0, st,
                           static_cast<const float*>(x), static_cast<const float*>(w), rows, N, K, P));
    VLG_LAUNCH_CHECK();
    return VLG_OK;
  }
  VLG_CUDA(launch_kernel(gemm_f32_kernel<EPI>, dim3((groups + kF32Warps - 1) / kF32Warps), dim3(kF32Warps * 32), 0, st,
                         static_cast<const float*>(x), static_cast<const float*>(w), rows, N, K, P));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

// ============================================================================ dispatch
// Tile shape per problem, by a small cost model measured on B200 (tools/gemm_phases.py): a CTA costs ~1.9 us of fixed
// work (barriers, TMEM, first TMA round trip, epilogue) plus, per 64-wide k-block, the time to pull its activation and
// weight tiles through the SM's L2 port (~116 GB/s); CTAs beyond one per SM run as a second wave.
template <int EPI>
static int launch_tc_auto(const void* x, const void* w, int rows, int N, int K, const EpiParams& P, cudaStream_t st) {
  static const int force_bn = [] { const char* e = getenv("VLG_GEMM_BN"); return e ? atoi(e) : 0; }();
  static const int force_bm = [] { const char* e = getenv("VLG_GEMM_BM"); return e ? atoi(e) : 0; }();
  const int sms = num_sms(), nkb = K / kBK;
  int best_bm = 128, best_bn = 16;
  double best_t = 1e30;
  const int bms[2] = {128, 64}, bns[4] = {16, 32, 64, 128};
  for (int bm : bms)
    for (int bn : bns) {
      if (N % bn) continue;
      if (force_bn && bn != force_bn) continue;
      if (force_bm && bm != force_bm) continue;
      if (rows <= 64 && bm == 128 && !force_bm) continue;            // never pad to 128 rows when 64 lanes suffice
      if (EPI == EPI_SWIGLU && bn < 16) continue;
      const long ctas = (long)(N / bn) * ((rows + bm - 1) / bm);
      const double waves = (double)((ctas + sms - 1) / sms);
      const double t = waves * (1.9 + nkb * (bm * 128.0 + bn * 128.0) / 116e3);
      if (t < best_t - 1e-9) { best_t = t; best_bm = bm; best_bn = bn; }
    }
#define VLG_TC_CASE(BM_, BN_) if (best_bm == BM_ && best_bn == BN_) return launch_tc<BM_, BN_, EPI>(x, w, rows, N, K, P, st)
  VLG_TC_CASE(128, 16); VLG_TC_CASE(128, 32); VLG_TC_CASE(128, 64); VLG_TC_CASE(128, 128);
  VLG_TC_CASE(64, 16); VLG_TC_CASE(64, 32); VLG_TC_CASE(64, 64); VLG_TC_CASE(64, 128);
#undef VLG_TC_CASE
  set_error("gemm(bf16): no tile shape for N=%d", N);
  return VLG_ERR_ARG;
}

int launch_gemm(int dtype, int epi, const void* x, const void* w, int rows, int N, int K, const EpiParams& P,
                cudaStream_t st) {
  VLG_REQUIRE(rows > 0 && N > 0 && K > 0, "gemm: bad shape rows=%d N=%d K=%d", rows, N, K);
  if (dtype == VLG_BF16) {
    VLG_REQUIRE(K % kBK == 0 && N % 16 == 0, "gemm(bf16): need K %% 64 == 0 and N %% 16 == 0 (K=%d N=%d)", K, N);
    switch (epi) {
      case EPI_STORE: return launch_tc_auto<EPI_STORE>(x, w, rows, N, K, P, st);
      case EPI_STORE_F32: return launch_tc_auto<EPI_STORE_F32>(x, w, rows, N, K, P, st);
      case EPI_RESID: return launch_tc_auto<EPI_RESID>(x, w, rows, N, K, P, st);
      case EPI_SWIGLU: return launch_tc_auto<EPI_SWIGLU>(x, w, rows, N, K, P, st);
      case EPI_QKV_ROPE: return launch_tc_auto<EPI_QKV_ROPE>(x, w, rows, N, K, P, st);
      case EPI_GELU: return launch_tc_auto<EPI_GELU>(x, w, rows, N, K, P, st);
    }
  } else if (dtype == VLG_F32) {
    VLG_REQUIRE(K % 4 == 0 && N % 8 == 0, "gemm(fp32): need K %% 4 == 0 and N %% 8 == 0 (K=%d N=%d)", K, N);
    switch (epi) {
      case EPI_STORE: case EPI_STORE_F32: return launch_f32<EPI_STORE>(x, w, rows, N, K, P, st);
      case EPI_RESID: return launch_f32<EPI_RESID>(x, w, rows, N, K, P, st);
      case EPI_SWIGLU: return launch_f32<EPI_SWIGLU>(x, w, rows, N, K, P, st);
      case EPI_QKV_ROPE: return launch_f32<EPI_QKV_ROPE>(x, w, rows, N, K, P, st);
      case EPI_GELU: return launch_f32<EPI_GELU>(x, w, rows, N, K, P, st);
    }
  }
  set_error("gemm: unsupported dtype %d / epilogue %d", dtype, epi);
  return VLG_ERR_ARG;
}

}  // namespace vlg

extern "C" int vlg_linear(const void* x_dev, const void* w_dev, void* out_dev, int rows, int N, int K, int dtype,
                          int out_f32, void* stream) {
  using namespace vlg;
  VLG_REQUIRE(x_dev && w_dev && out_dev, "vlg_linear: null pointer");
  EpiParams P{};
  P.out = out_dev;
  P.ld_out = N;
  return launch_gemm(dtype, out_f32 ? EPI_STORE_F32 : EPI_STORE, x_dev, w_dev, rows, N, K, P, (cudaStream_t)stream);
}

// microbenchmark instrumentation: phase timestamps (ns) of CTA 0 of the last tcgen05 GEMM run with VLG_GEMM_DEBUG & 4
extern "C" int vlg_debug_gemm_timestamps(unsigned long long* out16_host) {
  VLG_CUDA(cudaMemcpyFromSymbol(out16_host, vlg::g_gemm_ts, sizeof(unsigned long long) * 16));
  return VLG_OK;
}
