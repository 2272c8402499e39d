// Microbenchmark: how fast can persistent CTAs pull L2-resident bytes into shared memory with bulk (TMA) copies, as a
// function of (a) how many CTAs pull at once, (b) whether they pull the same bytes (activation-like) or their own
// (weight-like), (c) cluster multicast of the shared bytes.  Decides between the per-SM port and the chip-level L2 output
// as the bound of the layer-tail GEMM phases, and whether multicast lifts it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/ingest_bench ingest_bench.cu && ./bin/ingest_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kStages = 8;
constexpr int kChunk = 16384;
constexpr int kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) { while (!mbar_try(b, parity)) {} }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// arrive on the same-named barrier of cluster CTA `cta`
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* b, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// mode 0: every CTA streams the same `region` bytes (shared);  mode 1: CTA b streams its own region (src + b * region);
// mode 2: half of the chunks shared, half own (a GEMM unit: activation tile + weight tile);
// CS > 1: the shared bytes are multicast inside clusters of CS CTAs (each CTA fetches 1/CS of every chunk for everyone).
template <int CS>
__global__ void __launch_bounds__(kThreads, 1) ingest(const uint8_t* src, size_t region, int chunks_per_cta, int mode, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* ring = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kChunk);
  uint64_t* empty = full + kStages;
  const uint32_t rank = CS > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) cluster_sync();
  const uint8_t* own = src + (size_t)(1 + blockIdx.x) * region;
  const int per_region = (int)(region / kChunk);
  if (threadIdx.x == 0) {                       // producer
    for (int i = 0; i < chunks_per_cta; ++i) {
      const int s = i % kStages, it = i / kStages;
      mbar_wait(&empty[s], (it & 1) ^ 1);
      mbar_expect_tx(&full[s], kChunk);
      const bool shared_chunk = mode == 0 || (mode == 2 && (i & 1) == 0);
      const size_t off = (size_t)(i % per_region) * kChunk;
      if (shared_chunk && CS > 1) {
        const uint32_t piece = kChunk / CS;
        bulk_load_mc(ring + s * kChunk + rank * piece, src + off + rank * piece, piece, &full[s], (uint16_t)((1u << CS) - 1));
      } else if (shared_chunk) {
        bulk_load(ring + s * kChunk, src + off, kChunk, &full[s]);
      } else {
        bulk_load(ring + s * kChunk, own + off, kChunk, &full[s]);
      }
    }
  } else if (threadIdx.x == 32) {               // consumer: waits for the bytes, frees the slot (in every cluster CTA)
    unsigned acc = 0;
    for (int i = 0; i < chunks_per_cta; ++i) {
      const int s = i % kStages, it = i / kStages;
      mbar_wait(&full[s], it & 1);
      acc += ring[s * kChunk + (i & 1023)];
      if (CS > 1) { for (uint32_t c = 0; c < CS; ++c) mbar_arrive_remote(&empty[s], c); }
      else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
    if (acc == 0xffffffffu) *sink = acc;
  }
  __syncthreads();
  if (CS > 1) cluster_sync();
}

template <int CS>
float run(int grid, const uint8_t* src, size_t region, int chunks, int mode, unsigned* sink) {
  const int smem = kStages * kChunk + 256;
  cudaFuncSetAttribute(ingest<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (CS > 1) {
    int nc = 0;
    cudaOccupancyMaxActiveClusters(&nc, ingest<CS>, &cfg);
    if (nc * CS < grid) { printf("    (cluster %d: only %d clusters = %d CTAs co-resident; grid %d)\n", CS, nc, nc * CS, grid); }
  }
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaLaunchKernelEx(&cfg, ingest<CS>, src, region, chunks, mode, sink);     // warm L2
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  cudaLaunchKernelEx(&cfg, ingest<CS>, src, region, chunks, mode, sink);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  if (e != cudaSuccess) printf("    CUDA error: %s\n", cudaGetErrorString(e));
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t region = 256 * 1024;                       // 256 KB: a [128 x 1024] bf16 activation slab
  uint8_t* src; cudaMalloc(&src, region * (sms + 2)); cudaMemset(src, 1, region * (sms + 2));   // 38 MB: L2-resident
  unsigned* sink; cudaMalloc(&sink, 4);
  const int chunks = 2048;                                // 32 MB per CTA
  const char* names[3] = {"shared bytes", "own bytes", "half shared, half own"};
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {1, 16, 44, 88, 128, 148}) {
      float ms = run<1>(grid, src, region, chunks, mode, sink);
      double per = (double)chunks * kChunk / (ms * 1e-3) / 1e9;
      printf("unicast   %-22s grid %3d: %7.1f GB/s per CTA, %6.2f TB/s into shared memory\n", names[mode], grid, per, per * grid / 1e3);
    }
  for (int mode : {0, 2}) {
    for (int grid : {16, 88, 148}) {
      float ms = run<2>(grid, src, region, chunks, mode, sink);
      double per = (double)chunks * kChunk / (ms * 1e-3) / 1e9;
      printf("cluster 2 %-22s grid %3d: %7.1f GB/s per CTA, %6.2f TB/s into shared memory\n", names[mode], grid, per, per * grid / 1e3);
    }
    for (int grid : {16, 88, 132, 144}) {
      float ms = run<4>(grid, src, region, chunks, mode, sink);
      double per = (double)chunks * kChunk / (ms * 1e-3) / 1e9;
      printf("cluster 4 %-22s grid %3d: %7.1f GB/s per CTA, %6.2f TB/s into shared memory\n", names[mode], grid, per, per * grid / 1e3);
    }
    for (int grid : {16, 88, 128}) {
      float ms = run<8>(grid, src, region, chunks, mode, sink);
      double per = (double)chunks * kChunk / (ms * 1e-3) / 1e9;
      printf("cluster 8 %-22s grid %3d: %7.1f GB/s per CTA, %6.2f TB/s into shared memory\n", names[mode], grid, per, per * grid / 1e3);
    }
  }
  return 0;
}
