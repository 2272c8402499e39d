// CTA-pair GEMM building block: out[R <= 128, N] = x[R, K] . w[N, K]^T in bf16 with fp32 accumulation, computed by clusters of
// two CTAs that share ONE tcgen05.mma.cta_group::2 per K step (M = 128 split 64 + 64 over the pair, N = BNP).
//
// Why pairs: at R = 128 every GEMM unit of the decode step has to pull the whole [128 x K_slice] activation slab through its
// SM's L2 port (110-128 GB/s per SM, tools/micro/ingest_bench.cu), and that -- not HBM, not the tensor pipe -- paces the
// w1|w3 and wqkv phases of the layer tail.  In a pair each CTA loads only ITS 64 activation rows and ITS half
// of the weight tile; the MMA reads both halves of both operands from the two shared memories, so the bytes through each
// SM's port drop from (128 + BN) to (64 + BNP / 2) rows per K block, and one issuing thread feeds two SMs' tensor cores.
//
//   rank r of a pair:  A tile  = x rows [64 r, 64 r + 64)           (TMA box 64 x 64, 128B swizzle)
//                      B tile  = w rows [n0 + r BNP/2, + BNP/2)
//                      D       = rows [64 r, 64 r + 64) x BNP columns in ITS tensor memory, the "2x2" layout of
//                                cute::UMMA::tmem_frg_2sm (cute/atom/mma_traits_sm100.hpp): column n < BNP/2 -> lane m,
//                                column n; column n >= BNP/2 -> lane 64 + m, column n - BNP/2.
//   barriers:          full[s]  lives in the leader (rank 0): armed there for the bytes of both CTAs; both producers' TMA loads
//                               (cp.async.bulk.tensor...cta_group::2) complete on it
//                      empty[s], acc_ready: one per CTA, arrived by tcgen05.commit...multicast::cluster from the leader
//
// This file is the verified reference of those mechanics (tests/test_gpu_ops.py::test_pair_gemm_*); the layer-tail kernel
// builds its pair phases from the same pieces.
#include "common.cuh"
#include "kernels.h"

namespace vlg {

constexpr int kPgThreads = 192;     // warp 0: TMA producer, warp 1: MMA issuer (leader) + TMEM owner, warps 2-5: TMEM readers
constexpr int kPgStages = 6;
constexpr int kPgABytes = 64 * 128;                 // 64 rows x 64 k
struct PgSmem {
  uint64_t full[kPgStages], empty[kPgStages];
  uint64_t acc_ready;
  uint32_t tmem_slot;
};

template <int BNP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPgThreads, 1)
pair_gemm_kernel(const __grid_constant__ TensorMap tmX, const __grid_constant__ TensorMap tmW, float* __restrict__ out, int rows,
                 int N, int K) {
  constexpr int kWBytes = (BNP / 2) * 128;
  constexpr int kStage = kPgABytes + kWBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  PgSmem& sm = *reinterpret_cast<PgSmem*>(ring + kPgStages * kStage);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n0 = (blockIdx.x >> 1) * BNP;
  const int nkb = K / 64;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < kPgStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.acc_ready, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_pair<BNP / 2 < 32 ? 32 : BNP / 2>(&sm.tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_acc = sm.tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kPgStages, it = kb / kPgStages;
        mbar_wait(&sm.empty[s], (it & 1) ^ 1);
        if (rank == 0) mbar_expect_tx(&sm.full[s], 2 * kStage);
        uint8_t* a = ring + s * kStage;
        tma_load_2d_pair(a, &tmX, &sm.full[s], kb * 64, (int)rank * 64, kEvictLast);
        tma_load_2d_pair(a + kPgABytes, &tmW, &sm.full[s], kb * 64, n0 + (int)rank * (BNP / 2), kEvictFirst);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BNP);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kPgStages, it = kb / kPgStages;
        mbar_wait(&sm.full[s], it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(ring + s * kStage);
        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + kPgABytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
        umma_commit_pair(&sm.empty[s]);
      }
      umma_commit_pair(&sm.acc_ready);
    }
    __syncwarp();
  } else {
    const int quad = warp & 3;                       // TMEM lanes [32 quad, 32 quad + 32)
    const int r = (int)rank * 64 + (quad & 1) * 32 + lane;
    const int col0 = (quad >> 1) * (BNP / 2);        // lanes 64.. hold the second half of the columns
    mbar_wait(&sm.acc_ready, 0);
    tc_fence_after();
    for (int c = 0; c < BNP / 2; c += 16) {
      float v[16];
      tmem_ld16(tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
      if (r < rows && n0 + col0 + c < N) {
#pragma unroll
        for (int i = 0; i < 16; ++i) out[(size_t)r * N + n0 + col0 + c + i] = v[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // both CTAs are done with the pair's tensor memory
  if (warp == 1) tmem_dealloc_pair<BNP / 2 < 32 ? 32 : BNP / 2>(tmem_acc);
}

}  // namespace vlg

// verification entry: out[rows, N] fp32 = x[rows, K] . w[N, K]^T, bf16 operands; rows <= 128, K % 64 == 0, N % bnp == 0
extern "C" int vlg_debug_pair_gemm(const void* x_dev, const void* w_dev, float* out_dev, int rows, int N, int K, int bnp, void* stream) {
  using namespace vlg;
  VLG_REQUIRE(x_dev && w_dev && out_dev, "vlg_debug_pair_gemm: null pointer");
  VLG_REQUIRE(rows > 0 && rows <= 128 && K > 0 && K % 64 == 0 && (bnp == 64 || bnp == 128) && N > 0 && N % bnp == 0,
              "vlg_debug_pair_gemm: rows %d (<= 128), K %d (%% 64), N %d (%% bnp), bnp %d (64 | 128)", rows, K, N, bnp);
  TensorMap tmX{}, tmW{};
  int rc;
  if ((rc = make_tmap_bf16_2d(&tmX, x_dev, (uint64_t)rows, (uint64_t)K, (uint64_t)K * 2, 64, 64, true))) return rc;
  if ((rc = make_tmap_bf16_2d(&tmW, w_dev, (uint64_t)N, (uint64_t)K, (uint64_t)K * 2, (uint32_t)bnp / 2, 64, true))) return rc;
  const dim3 grid(2 * (N / bnp));
  cudaStream_t st = (cudaStream_t)stream;
  if (bnp == 128) {
    const int smem = kPgStages * (kPgABytes + 64 * 128) + 1024 + 256;
    VLG_CUDA(cudaFuncSetAttribute(pair_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pair_gemm_kernel<128><<<grid, kPgThreads, smem, st>>>(tmX, tmW, out_dev, rows, N, K);
  } else {
    const int smem = kPgStages * (kPgABytes + 32 * 128) + 1024 + 256;
    VLG_CUDA(cudaFuncSetAttribute(pair_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pair_gemm_kernel<64><<<grid, kPgThreads, smem, st>>>(tmX, tmW, out_dev, rows, N, K);
  }
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}
