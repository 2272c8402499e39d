// Microbenchmark: what does one thread pay per tcgen05.mma (M = 128, N = 64, K = 16, SS) in a TMA-ring style issue loop?
// Separates the instruction itself from the per-k-block overheads of the layer-tail kernels (mbarrier try_wait on an already
// completed phase, tcgen05.commit per k-block) and from the pressure other warps put on the barrier unit by spinning on an
// mbarrier with all 32 lanes (what the TMEM-reader warps do while the main loop runs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/mma_issue_bench mma_issue_bench.cu && ./bin/mma_issue_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// A operand from tensor memory (the activation tile copied there by tcgen05.cp first)
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 128 rows x 256 bits (one K = 16 slice of a K-major tile) shared -> tensor memory
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

constexpr int kStages = 6;
struct Sm {
  uint64_t full[kStages], empty[kStages], done, never;
  uint32_t tmem;
};

// mode bits: 1 = try_wait on a completed barrier per k-block, 2 = a second one, 4 = commit per k-block,
//            8 = 8 bystander warps spin on `never` with all lanes, 16 = the same with one lane per warp
__global__ void __launch_bounds__(320, 1) bench(int mode, int m, int n, int nkb, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  Sm& sm = *reinterpret_cast<Sm*>(tiles + kStages * 32768);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.done, 1);
    mbar_init(&sm.never, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < kStages; ++s) mbar_arrive(&sm.full[s]);        // phase 0 of every `full` is complete: waits on parity 0 pass
  }
  for (int i = threadIdx.x; i < kStages * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x3c003c00u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sm.tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t acc = sm.tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  if (warp == 1 && lane == 0) {
    long long t0 = clock64();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      if (mode & 1) mbar_wait(&sm.full[s], 0);
      if (mode & 2) mbar_wait(&sm.full[(s + 1) % kStages], 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a = smem_u32(tiles + s * 32768);
      const uint64_t ad = desc_sw128(a), bd = desc_sw128(n > 128 ? smem_u32(tiles) : a + 16384);
      if (mode & 32) {          // TS: copy the 4 K-slices into a double-buffered A region of tensor memory, then MMA from there
        const uint32_t abuf = acc + 128 + 32 * (kb & 1);
#pragma unroll
        for (int k = 0; k < 4; ++k) tmem_cp_128x256b(abuf + 8 * k, ad + 2 * k);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ts(acc, abuf + 8 * k, bd + 2 * k, idesc, (kb | k) != 0);
      } else if (mode & 64) {   // two accumulators alternating: is it the dependent accumulation that costs?
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(acc + 64 * (k & 1), ad + 2 * k, bd + 2 * k, idesc, (kb | (k >> 1)) != 0);
      } else if (mode & 128) {  // four accumulators
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(acc + 64 * k, ad + 2 * k, bd + 2 * k, idesc, kb != 0);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
      }
      if (mode & 4) commit(&sm.empty[s]);
    }
    commit(&sm.done);
    long long t1 = clock64();
    mbar_wait(&sm.done, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    mbar_arrive(&sm.never);
  } else if (warp == 2 && lane == 0 && (mode & 256)) {
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      const uint32_t a = smem_u32(tiles + s * 32768);
      const uint64_t ad = desc_sw128(a), bd = desc_sw128(a + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma(acc + 128, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
    }
    commit(&sm.never);
  } else if (warp >= 3 && (mode & 24)) {
    if (mode & 8) mbar_wait(&sm.never, 0);
    else { if (lane == 0) mbar_wait(&sm.never, 0); __syncwarp(); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(acc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// The whole warp walks the loop (warp-uniform control flow and operands), only the instruction itself is issued by one lane.
// variant 0: elect once around the 4 MMAs of a k-block; 1: elect per instruction; 2: no runtime stage arithmetic (unrolled ring)
__global__ void __launch_bounds__(320, 1) bench_uniform(int variant, int m, int n, int nkb, long long* out) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  Sm& sm = *reinterpret_cast<Sm*>(tiles + kStages * 32768);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int s = 0; s < kStages; ++s) mbar_arrive(&sm.full[s]);
  }
  for (int i = threadIdx.x; i < kStages * 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x3c003c00u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&sm.tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t acc = sm.tmem;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  if (warp == 1) {
    const uint32_t base = smem_u32(tiles);
    long long t0 = clock64();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kStages;
      if (variant & 4) mbar_wait(&sm.full[s], 0);
      const uint32_t a = base + s * 32768;
      const uint64_t ad = desc_sw128(a), bd = desc_sw128(a + 16384);
      if (variant & 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (elect_one()) umma(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
      } else {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        }
      }
      if (variant & 2) { if (elect_one()) commit(&sm.empty[s]); }
      __syncwarp();
    }
    if (elect_one()) commit(&sm.done);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&sm.done, 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(acc) : "memory");
}

int main() {
  const int smem = kStages * 32768 + 2048;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out; cudaMallocManaged(&out, 16);
  const int nkb = 64;
  cudaFuncSetAttribute(bench_uniform, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int m : {64, 128})
    for (int n : {16, 64, 128, 256})
      for (int variant : {0, 1, 2, 6}) {
        for (int rep = 0; rep < 2; ++rep) { bench_uniform<<<1, 320, smem>>>(variant, m, n, nkb, out); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError();
        printf("warp-uniform loop M=%3d N=%3d elect %s commit/kb=%d wait=%d : issue %6.1f, complete %6.1f cycles per MMA  %s\n", m, n, (variant & 1) ? "per MMA" : "per kb ",
               (variant & 2) ? 1 : 0, (variant & 4) ? 1 : 0, out[0] / (4.0 * nkb), out[1] / (4.0 * nkb), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  const int modes[] = {0, 64, 128, 256};
  for (int m : {64, 128})
    for (int n : {16, 32, 64, 128, 256})
      for (int mode : modes) {
        if (n > 64 && mode != 0) continue;
        for (int rep = 0; rep < 2; ++rep) { bench<<<1, 320, smem>>>(mode, m, n, nkb, out); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError();
        printf("M=%3d N=%3d %s wait%s%s commit/kb=%d : issue %6.1f cycles per MMA, complete %6.1f cycles per MMA  %s\n", m, n, (mode & 64) ? "2 accumulators" : (mode & 128) ? "4 accumulators" : (mode & 256) ? "2 threads     " : "1 accumulator ",
               (mode & 1) ? "1" : "0", (mode & 2) ? "+1" : "  ", (mode & 4) ? 1 : 0, out[0] / (4.0 * nkb), out[1] / (4.0 * nkb), e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
