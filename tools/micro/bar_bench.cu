// Microbenchmark: latency of grid-barrier variants on co-resident persistent CTAs (one per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bar_bench bar_bench.cu && ./bar_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int V>
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned seen;
    if (V == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
      __threadfence();
    } else if (V == 1) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else if (V == 2) {
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
    } else if (V == 3) {   // relaxed everything (NOT a correct barrier for data; lower bound of the atomic round trips)
      asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
    } else if (V == 4) {   // returning atomic: the last arriver knows immediately
      unsigned old;
      asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
      if (old + 1 < target) do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else if (V == 5) {   // volatile poll
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
  }
  __syncthreads();
}

// stores: bytes each of the first 128 threads writes (coalesced 32B pieces) before every barrier
template <int V>
__global__ void __launch_bounds__(256, 1) k(unsigned* ctr, float* buf, int iters, int store_floats) {
  unsigned target = 0;
  float* mine = buf + ((size_t)blockIdx.x * 128 + (threadIdx.x & 127)) * 128;
  for (int i = 0; i < iters; ++i) {
    if (threadIdx.x < 128)
      for (int j = 0; j < store_floats; j += 8)
        asm volatile("st.global.cg.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(mine + j), "f"((float)i) : "memory");
    grid_barrier<V>(ctr, target);
  }
}

template <int V> void run(const char* name, int sms, unsigned* ctr, float* buf) {
  for (int sf : {0, 64, 128}) {
    cudaMemset(ctr, 0, 4);
    const int iters = 2000;
    k<V><<<sms, 256>>>(ctr, buf, 10, sf);
    cudaMemset(ctr, 0, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<V><<<sms, 256>>>(ctr, buf, iters, sf);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-28s stores %4d B/thread: %.3f us per barrier (%s)\n", name, sf * 4, ms * 1e3 / iters, cudaGetErrorString(e));
  }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* ctr; cudaMalloc(&ctr, 256);
  float* buf; cudaMalloc(&buf, (size_t)sms * 128 * 128 * 4);
  run<0>("V0 fence+red.rel+ld.acq+fence", sms, ctr, buf);
  run<1>("V1 red.rel+ld.relaxed+fence", sms, ctr, buf);
  run<2>("V2 red.rel+ld.acq", sms, ctr, buf);
  run<3>("V3 all relaxed (bound)", sms, ctr, buf);
  run<4>("V4 atom.rel ret+relaxed+fence", sms, ctr, buf);
  run<5>("V5 red.rel+ld.volatile+fence", sms, ctr, buf);
  return 0;
}
