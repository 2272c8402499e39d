// RMSNorm and embedding-row gathers of the decode step.
#include "common.cuh"
#include "kernels.h"

namespace vlg {

// ---------------------------------------------------------------- RMSNorm (gpt.py:143-148)
// out = T( T(x * rsqrt(mean(x^2) + eps)) * w ): normalise in fp32, round to the storage type, THEN scale by
// the weight and round again -- the reference's two roundings, not a fused single rounding.
constexpr int kNormThreads = 128;
constexpr int kNormMaxVec = 4;  // 8-element vectors per thread -> dim <= 4096

template <typename T>
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ out, int dim, float eps) {
  __shared__ float s_part[kNormThreads / 32];
  const int r = blockIdx.x, nvec = dim / 8;
  const T* xr = x + (size_t)r * dim;
  pdl_launch_dependents();
  // x depends on the previous kernel; the norm weight does not: fetch it before waiting
  float g[kNormMaxVec][8];
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    int vi = threadIdx.x + i * kNormThreads;
    if (vi < nvec) load8<T>(w + vi * 8, g[i]);
  }
  pdl_wait();
  float v[kNormMaxVec][8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    int vi = threadIdx.x + i * kNormThreads;
    if (vi < nvec) {
      load8<T>(xr + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = __fadd_rn(ss, __fmul_rn(v[i][j], v[i][j]));  // x*x rounded, then summed
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kNormThreads / 32; ++i) tot += s_part[i];
  const float rs = rsqrtf(__fdiv_rn(tot, (float)dim) + eps);
#pragma unroll
  for (int i = 0; i < kNormMaxVec; ++i) {
    int vi = threadIdx.x + i * kNormThreads;
    if (vi < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __fmul_rn(rnd<T>(__fmul_rn(v[i][j], rs)), g[i][j]);
      store8<T>(out + (size_t)r * dim + vi * 8, o);
    }
  }
}

int launch_rmsnorm(int dtype, const void* x, const void* w, void* out, int rows, int dim, float eps, cudaStream_t st) {
  VLG_REQUIRE(dim % 8 == 0 && dim <= 8 * kNormThreads * kNormMaxVec, "rmsnorm: dim %d unsupported", dim);
  if (rows == 0) return VLG_OK;
  if (dtype == VLG_BF16)
    VLG_CUDA(launch_kernel(rmsnorm_kernel<__nv_bfloat16>, dim3(rows), dim3(kNormThreads), 0, st, (const __nv_bfloat16*)x,
                           (const __nv_bfloat16*)w, (__nv_bfloat16*)out, dim, eps));
  else
    VLG_CUDA(launch_kernel(rmsnorm_kernel<float>, dim3(rows), dim3(kNormThreads), 0, st, (const float*)x, (const float*)w,
                           (float*)out, dim, eps));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

// ---------------------------------------------------------------- embedding gathers (gpt.py:78-83, 354)
// Row r of `out` = table[index(r)].  The index source covers the three callers:
//   idx64 != null : index = idx64[r]                                      (forward_pos API)
//   seq   != null : index = seq[(r % B) * seq_stride + *col_ptr + col_off] (decode loop: cat([x, x]), generate.py:92)
//   labels!= null : index = r < B ? labels[r] : null_class                (c2i prefill rows, generate.py:131-132)
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ table, T* __restrict__ out, int rows, int dim,
                                   const int64_t* __restrict__ idx64, const int32_t* __restrict__ seq, int seq_stride,
                                   const int* __restrict__ col_ptr, int col_off, const int64_t* __restrict__ labels,
                                   int B, int null_class, int table_rows) {
  const int r = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  int64_t id;
  if (idx64) id = idx64[r];
  else if (seq) id = seq[(size_t)(r % B) * seq_stride + (*col_ptr + col_off)];
  else id = r < B ? labels[r] : null_class;
  if (id < 0 || id >= table_rows) id = 0;  // out-of-range ids would fault in the reference; clamp instead
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * dim);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)r * dim);
  const int n16 = dim * (int)sizeof(T) / 16;
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}

int launch_gather(int dtype, const void* table, void* out, int rows, int dim, const int64_t* idx64,
                  const int32_t* seq, int seq_stride, const int* col_ptr, int col_off, const int64_t* labels, int B,
                  int null_class, int table_rows, cudaStream_t st) {
  if (rows == 0) return VLG_OK;
  if (dtype == VLG_BF16)
    VLG_CUDA(launch_kernel(gather_rows_kernel<__nv_bfloat16>, dim3(rows), dim3(128), 0, st, (const __nv_bfloat16*)table,
                           (__nv_bfloat16*)out, rows, dim, idx64, seq, seq_stride, col_ptr, col_off, labels, B,
                           null_class, table_rows));
  else
    VLG_CUDA(launch_kernel(gather_rows_kernel<float>, dim3(rows), dim3(128), 0, st, (const float*)table, (float*)out,
                           rows, dim, idx64, seq, seq_stride, col_ptr, col_off, labels, B, null_class, table_rows));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg

extern "C" int vlg_rmsnorm(const void* x_dev, const void* w_dev, void* out_dev, int rows, int dim, float eps, int dtype,
                           void* stream) {
  using namespace vlg;
  VLG_REQUIRE(x_dev && w_dev && out_dev, "vlg_rmsnorm: null pointer");
  VLG_REQUIRE(dtype == VLG_F32 || dtype == VLG_BF16, "vlg_rmsnorm: bad dtype %d", dtype);
  return launch_rmsnorm(dtype, x_dev, w_dev, out_dev, rows, dim, eps, (cudaStream_t)stream);
}
