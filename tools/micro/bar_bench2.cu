// Microbenchmark 2: grid-barrier variants that keep the pollers off the arrival counter's L2 line.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/bar_bench2 bar_bench2.cu && ./bin/bar_bench2
// V2  = the layer-tail kernel's barrier today: red.release on one word, every CTA polls that word (ld.acquire).
// V6  = returning atomic; the last arriver publishes the epoch to G group flags (one 128-byte line each), CTA b polls
//       flag b % G.  The arrival atomics no longer queue behind 148 pollers on the same L2 slice.
// V7  = V6 with the flags written by one warp of the last arriver (G = gridDim.x: a private line per CTA).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 384;

template <int V, int G>
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned* flags, unsigned& target, unsigned& epoch) {
  target += gridDim.x;
  epoch += 1;
  __syncthreads();
  if (V == 2) {
    if (threadIdx.x == 0) {
      unsigned seen;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory"); } while (seen < target);
    }
  } else if (V == 6) {
    if (threadIdx.x == 0) {
      unsigned old, seen;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
      if (old + 1 == target) {
        for (int g = 0; g < G; ++g) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + g * 32), "r"(epoch) : "memory");
      } else {
        const unsigned* f = flags + (blockIdx.x % G) * 32;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory"); } while (seen < epoch);
      }
    }
  } else if (V == 7) {
    __shared__ unsigned s_last;
    if (threadIdx.x == 0) {
      unsigned old;
      asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(ctr) : "memory");
      s_last = (old + 1 == target);
    }
    __syncthreads();
    if (s_last) {
      if (threadIdx.x < gridDim.x) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + threadIdx.x * 32), "r"(epoch) : "memory");
    } else if (threadIdx.x == 0) {
      unsigned seen;
      const unsigned* f = flags + blockIdx.x * 32;
      do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory"); } while (seen < epoch);
    }
  }
  __syncthreads();
}

template <int V, int G>
__global__ void __launch_bounds__(kThreads, 1) k(unsigned* ctr, unsigned* flags, float* buf, int iters, int store_floats) {
  unsigned target = 0, epoch = 0;
  float* mine = buf + ((size_t)blockIdx.x * 128 + (threadIdx.x & 127)) * 128;
  for (int i = 0; i < iters; ++i) {
    if (threadIdx.x < 128)
      for (int j = 0; j < store_floats; j += 8)
        asm volatile("st.global.cg.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(mine + j), "f"((float)i) : "memory");
    grid_barrier<V, G>(ctr, flags, target, epoch);
  }
}

template <int V, int G> void run(const char* name, int sms, unsigned* ctr, unsigned* flags, float* buf) {
  for (int sf : {0, 64}) {
    const int iters = 2000;
    cudaMemset(ctr, 0, 4); cudaMemset(flags, 0, 256 * 128);
    k<V, G><<<sms, kThreads>>>(ctr, flags, buf, 10, sf);
    cudaDeviceSynchronize();
    cudaMemset(ctr, 0, 4); cudaMemset(flags, 0, 256 * 128);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<V, G><<<sms, kThreads>>>(ctr, flags, buf, iters, sf);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-40s stores %4d B/thread: %.3f us per barrier (%s)\n", name, sf * 4, ms * 1e3 / iters, cudaGetErrorString(e));
  }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned *ctr, *flags; cudaMalloc(&ctr, 256); cudaMalloc(&flags, 256 * 128);
  float* buf; cudaMalloc(&buf, (size_t)sms * 128 * 128 * 4);
  run<2, 1>("V2 red.rel + ld.acq on the counter", sms, ctr, flags, buf);
  run<6, 1>("V6 atom + 1 flag line", sms, ctr, flags, buf);
  run<6, 4>("V6 atom + 4 flag lines", sms, ctr, flags, buf);
  run<6, 16>("V6 atom + 16 flag lines", sms, ctr, flags, buf);
  run<7, 0>("V7 atom + private flag per CTA (warp)", sms, ctr, flags, buf);
  return 0;
}
