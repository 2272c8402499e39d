// The layer tail (everything between two attentions of the bf16 decode step, see layer_tail.cu) for 1-2 rows: a single
// image with guidance, the t2v loop at batch 1 (BASELINE cfg-4, B = 1).
//
// At that size the tensor-core kernel is pure latency -- its six phases, five grid barriers and split-K exchanges take 25 us
// whatever the row count, for 6 us worth of weights -- and a matrix-vector product needs no tensor core.  Here a warp owns PAIRS
// of output columns (a RoPE pair, a (w1 x, w3 x) pair) and streams their weight rows straight from global memory into
// registers (16-byte loads, the next pair's rows in flight while this pair is multiplied), against activation rows it keeps
// packed in registers; every GEMV owns all of K, so there are no partials and no row phases: the RMSNorm in front of w1|w3 and
// wqkv is recomputed by every warp (2 rows x dim values) on the way in.  Three grid barriers instead of five:
//
//   wo + residual | B | RMSNorm -> w1|w3 -> SwiGLU | B | w2 + residual | B | RMSNorm -> wqkv -> RoPE + q / KV-cache write
//
// Rounding points are those of the per-op kernels (epilogue.cuh / elementwise.cu): GEMV results round to bf16
// before the elementwise tail, RMSNorm rounds before and after the weight multiply, RoPE products round separately.
#include "lt_common.cuh"

namespace vlg {

constexpr int kGvThreads = 256;
constexpr int kGvWarps = kGvThreads / 32;
constexpr int kGvMaxRows = 2;
// ND (template parameter of the kernel) = dim / 256: 256-element chunks per lane and row held in registers at a time -- K = dim in
// one go, K = hidden (w2) in segments of that many chunks.  Exact-size arrays: with a common maximum the compiler has to
// keep registers for chunks a model does not have (1.2 KB of spills at 2 rows).

struct GvParams {
  int rows, dim, hidden, has_front, has_qkv;
  const __nv_bfloat16 *attn, *wo, *w13, *w2, *wqkv, *norm1_w, *norm2_w;
  __nv_bfloat16 *h, *xn, *ffn;
  float eps;
  EpiParams rope;
  unsigned* bar;             // [0] barrier arrivals, [1] exit arrivals (the layer-tail counter region)
  volatile unsigned* err;
};

__device__ __forceinline__ void gv_grid_barrier(unsigned* ctr, unsigned& target, volatile unsigned* err) {
  target += gridDim.x;
  __syncthreads();
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
  __syncthreads();
}

__device__ __forceinline__ void gv_unpack(const uint4& v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}

// this lane's share of R activation rows of length K (K % 256 == 0): chunk ch = elements [(ch * 32 + lane) * 8, + 8), packed bf16
// (chunks [c0, c0 + MAXCH) of the row)
template <int R, int MAXCH>
__device__ __forceinline__ void gv_load_rows(const __nv_bfloat16* base, int K, int rows, int lane, uint4 (&a)[R][MAXCH], int c0 = 0) {
  const int nch = K >> 8;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch)
      if (c0 + ch < nch) a[r][ch] = ldg128_cg(base + (size_t)min(r, rows - 1) * K + ((c0 + ch) * 32 + lane) * 8);
}

// RMSNorm of the rows held in `a` (K = dim), in place: T(T(h * rsqrt(mean(h^2) + eps)) * w)   (gpt.py:143-148)
template <int R, int ND>
__device__ __forceinline__ void gv_norm_rows(const __nv_bfloat16* w, int dim, float eps, int lane, uint4 (&a)[R][ND]) {
  constexpr int nch = ND;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float ss = 0.f;
#pragma unroll
    for (int ch = 0; ch < ND; ++ch)
      if (ch < nch) {
        float f[8];
        gv_unpack(a[r][ch], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = __fadd_rn(ss, __fmul_rn(f[i], f[i]));
      }
    ss = warp_sum(ss);
    const float rs = rsqrtf(__fdiv_rn(ss, (float)dim) + eps);
#pragma unroll
    for (int ch = 0; ch < ND; ++ch)
      if (ch < nch) {
        float f[8], g[8];
        gv_unpack(a[r][ch], f);
        gv_unpack(__ldg(reinterpret_cast<const uint4*>(w + (ch * 32 + lane) * 8)), g);
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float y0 = __fmul_rn(f[2 * i], rs), y1 = __fmul_rn(f[2 * i + 1], rs);
          rnd2_bf16(y0, y1);
          o[i] = pack_bf16(__fmul_rn(y0, g[2 * i]), __fmul_rn(y1, g[2 * i + 1]));
        }
        a[r][ch] = make_uint4(o[0], o[1], o[2], o[3]);
      }
  }
}

// the weight rows of one column pair, this lane's chunks
template <int MAXCH> struct GvPair { uint4 w[2][MAXCH]; };

template <int MAXCH>
__device__ __forceinline__ void gv_load_pair(const __nv_bfloat16* W, int K, int n, int lane, GvPair<MAXCH>& p, int c0 = 0) {
  const int nch = K >> 8;
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int ch = 0; ch < MAXCH; ++ch)
      if (c0 + ch < nch) p.w[c][ch] = ldg_stream(W + (size_t)(n + c) * K + ((c0 + ch) * 32 + lane) * 8);
}

// acc[r][c] += sum over this lane's elements of chunks [c0, c0 + MAXCH) of a[r][k] * W[n + c][k]   (fp32, one chain per (row, column))
template <int R, int MAXCH>
__device__ __forceinline__ void gv_dot_accum(const GvPair<MAXCH>& p, const uint4 (&a)[R][MAXCH], int K, float (&acc)[R][2], int c0 = 0) {
  const int nch = K >> 8;
#pragma unroll
  for (int ch = 0; ch < MAXCH; ++ch)
    if (c0 + ch < nch) {
      float w0[8], w1[8];
      gv_unpack(p.w[0][ch], w0);
      gv_unpack(p.w[1][ch], w1);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float f[8];
        gv_unpack(a[r][ch], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[r][0] = fmaf(f[i], w0[i], acc[r][0]);
          acc[r][1] = fmaf(f[i], w1[i], acc[r][1]);
        }
      }
    }
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

// lane r finalises row r of the pair (n, n + 1)
template <int MODE, int R>
__device__ __forceinline__ void gv_finish(const GvParams& P, int n, const float (&acc)[R][2], int lane) {
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
    uint32_t* hp = reinterpret_cast<uint32_t*>(P.h + (size_t)r * P.dim + n);
    uint32_t hv;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(hv) : "l"(hp));
    const uint32_t o = pack_bf16(__fadd_rn(bf16lo(hv), a0), __fadd_rn(bf16hi(hv), a1));
    asm volatile("st.global.cg.u32 [%0], %1;" ::"l"(hp), "r"(o) : "memory");
  } else if constexpr (MODE == GV_SWIGLU) {            // gpt.py:167, w1 / w3 rows interleaved
    const float s = rnd<__nv_bfloat16>(silu_f(a0));
    P.ffn[(size_t)r * P.hidden + (n >> 1)] = __float2bfloat16_rn(__fmul_rn(s, a1));
  } else {                                             // gpt.py:215-227, 423-433, 177-185
    const EpiParams& E = P.rope;
    const int pos = E.pos_ptr ? *E.pos_ptr : E.pos_imm;
    const int sec = n / E.dim, cc = n - sec * E.dim, head = cc / E.D, dd = cc - head * E.D;
    float o0 = a0, o1 = a1;
    if (sec < 2) {
      const float2 cs = reinterpret_cast<const float2*>(E.rope)[(size_t)pos * (E.D >> 1) + (dd >> 1)];
      o0 = __fsub_rn(__fmul_rn(a0, cs.x), __fmul_rn(a1, cs.y));
      o1 = __fadd_rn(__fmul_rn(a1, cs.x), __fmul_rn(a0, cs.y));
    }
    __nv_bfloat16* dst;
    if (sec == 0) dst = static_cast<__nv_bfloat16*>(E.q_out) + (size_t)r * E.dim + cc;
    else dst = static_cast<__nv_bfloat16*>(sec == 1 ? E.k_cache : E.v_cache) + (((size_t)r * E.H + head) * E.S_max + pos) * E.D + dd;
    *reinterpret_cast<uint32_t*>(dst) = pack_bf16(o0, o1);
  }
}

// one GEMV phase with K = dim: the warps of the grid take the column pairs round-robin against the rows in `a`; `first` holds
// this warp's first pair, requested before the barrier in front of the phase (weights do not depend on it)
template <int MODE, int R, int ND>
__device__ __forceinline__ void gv_phase(const GvParams& P, const __nv_bfloat16* W, int N, const uint4 (&a)[R][ND],
                                         const GvPair<ND>& first, int gw, int tw, int lane) {
  const int K = P.dim;
  GvPair<ND> cur = first;
  for (int n = 2 * gw; n < N; n += 2 * tw) {
    GvPair<ND> nxt;
    const int n2 = n + 2 * tw;
    if (n2 < N) gv_load_pair<ND>(W, K, n2, lane, nxt);
    float acc[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = 0.f;
    gv_dot_accum<R, ND>(cur, a, K, acc);
    gv_reduce<R>(acc);
    gv_finish<MODE, R>(P, n, acc, lane);
    if (n2 < N) cur = nxt;
  }
}

// w2 (K = hidden, up to 22 chunks per lane): a pair's rows and the activation rows pass through the registers in segments of
// ND chunks; `first` holds segment 0 of this warp's first pair
template <int R, int ND>
__device__ __forceinline__ void gv_phase_w2(const GvParams& P, const GvPair<ND>& first, int gw, int tw, int lane) {
  const int K = P.hidden, nch = K >> 8;
  for (int n = 2 * gw; n < P.dim; n += 2 * tw) {
    float acc[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = 0.f;
    for (int c0 = 0; c0 < nch; c0 += ND) {
      GvPair<ND> wseg;
      if (c0 == 0 && n == 2 * gw) wseg = first;
      else gv_load_pair<ND>(P.w2, K, n, lane, wseg, c0);
      uint4 a[R][ND];
      gv_load_rows<R, ND>(P.ffn, K, P.rows, lane, a, c0);
      gv_dot_accum<R, ND>(wseg, a, K, acc, c0);
    }
    gv_reduce<R>(acc);
    gv_finish<GV_RESID, R>(P, n, acc, lane);
  }
}

template <int R, int ND>
__global__ void __launch_bounds__(kGvThreads, 1) layer_tail_gemv_kernel(const __grid_constant__ GvParams P) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kGvWarps + warp, tw = gridDim.x * kGvWarps;
  const int d = P.dim, hid = P.hidden;
  unsigned bar_target = 0;
  pdl_launch_dependents();
  // the first phase's weights do not depend on the previous kernel
  GvPair<ND> fd;
  if (P.has_front) { if (2 * gw < d) gv_load_pair<ND>(P.wo, d, 2 * gw, lane, fd); }
  else if (2 * gw < 3 * d) gv_load_pair<ND>(P.wqkv, d, 2 * gw, lane, fd);
  pdl_wait();

  if (P.has_front) {
    {
      uint4 a[R][ND];
      gv_load_rows<R, ND>(P.attn, d, P.rows, lane, a);
      gv_phase<GV_RESID, R, ND>(P, P.wo, d, a, fd, gw, tw, lane);
    }
    if (2 * gw < 2 * hid) gv_load_pair<ND>(P.w13, d, 2 * gw, lane, fd);
    gv_grid_barrier(P.bar, bar_target, P.err);
    {
      uint4 a[R][ND];
      gv_load_rows<R, ND>(P.h, d, P.rows, lane, a);
      gv_norm_rows<R, ND>(P.norm1_w, d, P.eps, lane, a);
      gv_phase<GV_SWIGLU, R, ND>(P, P.w13, 2 * hid, a, fd, gw, tw, lane);
    }
    if (2 * gw < d) gv_load_pair<ND>(P.w2, hid, 2 * gw, lane, fd);
    gv_grid_barrier(P.bar, bar_target, P.err);
    gv_phase_w2<R, ND>(P, fd, gw, tw, lane);
    if (P.has_qkv && 2 * gw < 3 * d) gv_load_pair<ND>(P.wqkv, d, 2 * gw, lane, fd);
    gv_grid_barrier(P.bar, bar_target, P.err);
  }
  if (P.has_qkv) {
    uint4 a[R][ND];
    gv_load_rows<R, ND>(P.h, d, P.rows, lane, a);
    gv_norm_rows<R, ND>(P.norm2_w, d, P.eps, lane, a);
    gv_phase<GV_ROPE, R, ND>(P, P.wqkv, 3 * d, a, fd, gw, tw, lane);
  } else if (gw == 0) {                                  // after the last layer: the head GEMM reads xn from global memory
    uint4 a[R][ND];
    gv_load_rows<R, ND>(P.h, d, P.rows, lane, a);
    gv_norm_rows<R, ND>(P.norm2_w, d, P.eps, lane, a);
    const int nch = d >> 8;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int ch = 0; ch < ND; ++ch)
        if (ch < nch && r < P.rows) *reinterpret_cast<uint4*>(P.xn + (size_t)r * d + (ch * 32 + lane) * 8) = a[r][ch];
  }

  // leave the counters at zero for the next launch
  __syncthreads();
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
  return !off && rows >= 1 && rows <= kGvMaxRows && dim % 256 == 0 && hidden % 256 == 0 &&
         (dim == 768 || dim == 1024 || dim == 1280 || dim == 1536);    // dim 2048 spills at 2 rows (255 registers): tensor-core tail
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
  const dim3 grid(num_sms()), block(kGvThreads);
#define VLG_GV_CASE(ND_)                                                                                 \
  if (a.dim == ND_ * 256) {                                                                              \
    if (a.rows == 1) VLG_CUDA(launch_kernel(layer_tail_gemv_kernel<1, ND_>, grid, block, 0, st, P));     \
    else VLG_CUDA(launch_kernel(layer_tail_gemv_kernel<2, ND_>, grid, block, 0, st, P));                 \
  }
  VLG_GV_CASE(3) VLG_GV_CASE(4) VLG_GV_CASE(5) VLG_GV_CASE(6)
#undef VLG_GV_CASE
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg
