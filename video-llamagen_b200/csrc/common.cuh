// Shared device/host helpers for the vlg kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vlg.h"

namespace vlg {

// ------------------------------------------------------------------ error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define VLG_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) return ::vlg::cuda_fail(_e, #expr, __FILE__, __LINE__);  \
  } while (0)

#define VLG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::vlg::set_error(__VA_ARGS__);      \
      return VLG_ERR_ARG;                 \
    }                                     \
  } while (0)

void count_launches(uint64_t n);
// after every kernel launch: account for it and surface launch-configuration errors
#define VLG_LAUNCH_CHECK()        \
  do {                            \
    ::vlg::count_launches(1);     \
    VLG_CUDA(cudaGetLastError()); \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// Every decode-step kernel is launched with programmatic stream serialization: it may start while its
// predecessor drains, runs its input-independent prologue (barrier init, TMEM alloc, descriptor prefetch, weight
// prefetch), and only then waits for the predecessor's memory with griddepcontrol.wait.  VLG_PDL=0 disables it.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same with thread-block clusters of `cluster` CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// wait until the preceding kernel's writes are visible (no-op when launched without the PDL attribute)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// allow the next kernel in the stream to begin launching
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ------------------------------------------------------------------ small device utils
constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Storage-type traits: the decode path runs in bf16 (product) or fp32 (exact greedy contract).
template <typename T> struct Num;
template <> struct Num<float> {
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
};
template <> struct Num<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// round-trip through the storage type: models "the library op returned a T tensor"
template <typename T> __device__ __forceinline__ float rnd(float v) { return Num<T>::to_f(Num<T>::from_f(v)); }

// two bf16 packed in a 32-bit word -> two floats (low half first)
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// round two floats to bf16 and back with ONE packed conversion (cvt.rn.bf16x2.f32 -> F2FP, ALU-rate) instead of two scalar
// F2F conversions, which share the 16-lane-per-clock pipe with MUFU: in the SwiGLU epilogue the scalar roundings alone were
// ~1000 cycles per CTA.  Same round-to-nearest-even as rnd<__nv_bfloat16>.
__device__ __forceinline__ void rnd2_bf16(float& a, float& b) {
  const uint32_t p = pack_bf16(a, b);
  a = bf16lo(p);
  b = bf16hi(p);
}

// 128-bit streaming load that does not allocate in L1 (data read exactly once)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ------------------------------------------------------------------ mbarrier / TMA / tcgen05 PTX
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug turns into a trap (reported as a CUDA error) instead of a hung GPU.  The report is a real
// function, not inlined: a printf call site is ~25 instructions, and the persistent kernels have dozens of waits whose code is
// fetched cold on every launch.
static __device__ __noinline__ void mbar_timeout() {
  printf("vlg: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) mbar_timeout();
  }
}

// 2D tiled TMA load: box at (c0 = inner coordinate, c1 = row coordinate) -> smem, completes on bar
__device__ __forceinline__ void tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// L2 eviction policies (createpolicy encodings used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 / TMEM
template <int kCols> __device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on bar when they retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// warp-collective: lane i reads TMEM lane (base_lane + i), 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- CTA pairs: tcgen05 cta_group::2 (verified by pair_gemm.cu / tests/test_gpu_ops.py::test_pair_gemm_cta_group_2)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the pair's leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols> __device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// 2D tiled TMA load of this CTA's tile into its own shared memory; the bytes are accounted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem of both]; issued by one thread of the leader
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far arrive on the barrier at this offset in BOTH CTAs when they retire
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 128 B, 8-row groups 1024 B apart).
// Field layout follows cute::UMMA::SmemDescriptor (cute/arch/mma_sm100_desc.hpp).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);        // start address   [0,14)
  d |= static_cast<uint64_t>(1) << 16;                            // LBO (unused for swizzled K-major) [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                    // SBO = 1024 B    [32,46)
  d |= static_cast<uint64_t>(1) << 46;                            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                            // layout = SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: A=B=bf16, D=fp32, both K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------ host: tensor maps
// Encoded through the driver entry point fetched at run time (no link-time libcuda dependency).
struct alignas(64) TensorMap { uint8_t bytes[128]; };
// 2D row-major bf16 tensor [rows, cols] (cols contiguous), box [box_rows, box_cols], 128B swizzle
int make_tmap_bf16_2d(TensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                      uint32_t box_rows, uint32_t box_cols, bool swizzle128);

int num_sms();                       // SM count of the current device
constexpr int kMaxDevices = 64;
int current_device();                // cudaGetDevice clamped to [0, kMaxDevices)
// cudaFuncSetAttribute is per device: run `set` once for every (call site, device)
struct PerDeviceOnce {
  bool done[kMaxDevices] = {};
  template <typename F> cudaError_t operator()(F&& set) {
    const int d = current_device();
    if (done[d]) return cudaSuccess;
    cudaError_t e = set();
    if (e == cudaSuccess) done[d] = true;
    return e;
  }
};

}  // namespace vlg
