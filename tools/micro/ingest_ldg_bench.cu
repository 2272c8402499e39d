// Microbenchmark: is the ~110 GB/s per SM that TMA (bulk copies) pulls out of L2 a limit of the copy engine or of the SM's
// port?  Streams an L2-resident region into one CTA per SM through (a) bulk copies, (b) 16-byte cp.async (LDGSTS) issued by
// all threads, (c) plain 16-byte loads into registers, (d) bulk copies and cp.async at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/ingest_ldg_bench ingest_ldg_bench.cu && ./bin/ingest_ldg_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kThreads = 512;
constexpr int kChunk = 16384;
constexpr int kStages = 6;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// mode 0: bulk copies (warp 0 thread 0), 1: cp.async by every thread, 2: ld.global.nc into registers, 3: bulk (first half of
// the chunks) + cp.async (second half) concurrently
__global__ void __launch_bounds__(kThreads, 1) ingest(const uint8_t* src, size_t region, int chunks, int mode, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;                                   // kStages chunks for bulk, kStages more for cp.async
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kStages * kChunk);
  const uint8_t* own = src + (size_t)blockIdx.x * region;
  const int per_region = (int)(region / kChunk);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned acc = 0;
  const bool do_bulk = mode == 0 || mode == 3, do_cp = mode == 1 || mode == 3;
  if (do_bulk && threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      // the "consumer" is this same thread: wait for a slot's previous fill, then refill it (no MMA reads the bytes)
      for (int i = 0; i < chunks; ++i) {
        const int s = i % kStages, it = i / kStages;
        if (it > 0) mbar_wait(&full[s], (it - 1) & 1);
        mbar_expect_tx(&full[s], kChunk);
        bulk_load(ring + s * kChunk, own + (size_t)(i % per_region) * kChunk, kChunk, &full[s]);
      }
      for (int s = 0; s < kStages && s < chunks; ++s) mbar_wait(&full[s], ((chunks - 1 - s) / kStages) & 1);
      acc += ring[0];
    }
  } else if (do_cp && (mode == 1 || threadIdx.x >= 32)) {
    const int t = mode == 1 ? threadIdx.x : threadIdx.x - 32, nt = mode == 1 ? kThreads : kThreads - 32;
    uint8_t* ring2 = ring + kStages * kChunk;
    for (int i = 0; i < chunks; ++i) {
      const int s = i % kStages;
      const uint8_t* p = own + (size_t)((i + per_region / 2) % per_region) * kChunk;
      for (int o = t * 16; o < kChunk; o += nt * 16) cp_async16(ring2 + s * kChunk + o, p + o);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    acc += ring2[threadIdx.x];
  } else if (mode == 2) {
    uint4 a = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < chunks; ++i) {
      const uint4* p = reinterpret_cast<const uint4*>(own + (size_t)(i % per_region) * kChunk);
#pragma unroll
      for (int o = 0; o < kChunk / 16 / kThreads; ++o) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + threadIdx.x + o * kThreads));
        a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
      }
    }
    acc += a.x ^ a.y ^ a.z ^ a.w;
  }
  if (acc == 0xfffffffeu) *sink = acc;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t region = 256 * 1024;
  uint8_t* src; cudaMalloc(&src, region * (sms + 1)); cudaMemset(src, 1, region * (sms + 1));    // 38 MB: L2-resident
  unsigned* sink; cudaMalloc(&sink, 4);
  const int smem = 2 * kStages * kChunk + 256;
  cudaFuncSetAttribute(ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int chunks = 1024;                                 // 16 MB per CTA and path
  const char* names[4] = {"bulk copies (TMA)", "cp.async 16 B, 512 threads", "ld.global 16 B into registers", "bulk + cp.async together"};
  for (int mode = 0; mode < 4; ++mode)
    for (int grid : {1, 16, 148}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      ingest<<<grid, kThreads, smem>>>(src, region, chunks, mode, sink);
      cudaDeviceSynchronize();
      cudaEventRecord(a);
      ingest<<<grid, kThreads, smem>>>(src, region, chunks, mode, sink);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize(); if (e == cudaSuccess) e = cudaGetLastError();
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      const double bytes = (double)chunks * kChunk * (mode == 3 ? 2 : 1);
      printf("%-32s grid %3d: %7.1f GB/s per CTA  %s\n", names[mode], grid, bytes / (ms * 1e-3) / 1e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
