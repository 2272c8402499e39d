// Fused classifier-free-guidance mix + temperature + top-k / top-p filtering + softmax + token pick.
//
// Replaces, per decode step, generate.py:95-99 (CFG mix), :16-54 (top_k_top_p_filtering) and :57-66
// (sample): ~12 eager kernels incl. a topk(k=2000) and (for top-p) a full sort, plus the host sync that
// torch.multinomial can force.  One CTA owns one sample row; the V mixed logits live in shared memory
// from the first read to the final pick, so HBM traffic is one read of the two logit rows.
//
//  * top-k threshold: the k-th largest value, found by a 32-step bitwise selection over order-preserving keys held
//    in registers (one block-wide count per bit); entries strictly below it are dropped, ties at the threshold all
//    survive -- `logits < topk(...)[..., -1]` semantics.
//  * top-p: in-shared-memory bitonic sort of (value desc, index asc) keys, block prefix sum of the sorted
//    softmax, keep entry i iff i == 0 or cumsum[i-1] <= top_p (the reference's shift-right-by-one rule).
//  * pick: argmax(probs) (greedy) or argmax(probs / Exp(1)) -- the single-draw form of torch.multinomial --
//    with the noise either supplied by the caller (parity tests) or drawn in-kernel from Philox4x32-10.
#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace vlg {

constexpr int kSampThreads = 1024;

__device__ __forceinline__ uint32_t order_key(float f) {  // ascending uint order == ascending float order
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ float exp1_from_bits(uint32_t x) {  // Exp(1) draw, u in (0,1)
  float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
  return -logf(u);
}

struct BlockRed {  // block-wide reductions over kSampThreads threads
  float* sf; int* si;
  __device__ float max(float v) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sf[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sf[threadIdx.x & 31];
    r = warp_max(r);
    return r;
  }
  __device__ float sum(float v) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sf[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = sf[threadIdx.x & 31];
    r = warp_sum(r);
    return r;
  }
};

__global__ void __launch_bounds__(kSampThreads, 1) cfg_sample_kernel(SampleArgs a) {
  extern __shared__ __align__(16) uint8_t smraw[];
  float* val = reinterpret_cast<float*>(smraw);                                         // [V]
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smraw + (((size_t)a.V * 4 + 15) & ~size_t(15)));  // [Vp2] top-p only
  __shared__ float s_redf[32];
  __shared__ int s_redi[32];
  __shared__ unsigned s_cnt[3];
  __shared__ unsigned long long s_cut;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = a.V;
  BlockRed red{s_redf, s_redi};
  pdl_launch_dependents();
  pdl_wait();

  const int step = a.step_ptr ? *a.step_ptr : 0;
  const bool two_halves = a.rows == 2 * a.B;
  bool guided = a.p.guided != 0;
  // decode_n_tokens iteration i produces token i+1; guidance stops when i > cfg_interval (generate.py:113-114)
  if (a.cfg_interval > -1 && step - 1 > a.cfg_interval) guided = false;

  // ---- 1. CFG mix + temperature (generate.py:95-99, :58)
  const float* lc = a.logits + (size_t)b * V;
  const float* lu = a.logits + (size_t)(b + a.B) * V;
  const float inv_t = 1.0f / fmaxf(a.p.temperature, 1e-5f);   // ATen multiplies by the reciprocal of a scalar divisor
  float tmax = -INFINITY;
  for (int i = tid; i < V; i += kSampThreads) {
    float c = lc[i], x;
    if (two_halves && guided) {
      float u = lu[i];
      x = __fadd_rn(u, __fmul_rn(__fsub_rn(c, u), a.p.cfg_scale));
    } else {
      x = c;
    }
    x = __fmul_rn(x, inv_t);
    val[i] = x;
    tmax = fmaxf(tmax, x);
  }
  const float mx = red.max(tmax);   // row maximum; top-k / top-p never remove it
  __syncthreads();

  // ---- 2. top-k threshold = k-th largest key, built bit by bit from the MSB: keep bit b iff at least k keys are >= the
  // candidate.  32 block-wide counts; the keys stay in registers (V <= 16 K), each count is 16 compares + one REDUX +
  // one shared atomic per warp + one barrier.
  if (a.p.top_k > 0 && a.p.top_k < V) {
    constexpr int kPer = 16;
    const bool in_regs = V <= kSampThreads * kPer;
    uint32_t keys[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const int i = tid + j * kSampThreads;
      keys[j] = (in_regs && i < V) ? order_key(val[i]) : 0u;
    }
    if (tid < 3) s_cnt[tid] = 0;
    __syncthreads();
    const unsigned k = (unsigned)a.p.top_k;
    unsigned prefix = 0;
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned cand = prefix | (1u << bit);
      unsigned cnt = 0;
      if (in_regs) {
#pragma unroll
        for (int j = 0; j < kPer; ++j) cnt += keys[j] >= cand ? 1u : 0u;
      } else {
        for (int i = tid; i < V; i += kSampThreads) cnt += order_key(val[i]) >= cand ? 1u : 0u;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      unsigned* slot = &s_cnt[bit % 3];
      if (lane == 0 && cnt) atomicAdd(slot, cnt);
      __syncthreads();
      if (*slot >= k) prefix = cand;
      if (tid == 0) s_cnt[(bit + 1) % 3] = 0;       // recycled two iterations from now (a barrier lies in between)
    }
    const unsigned kth = prefix;
    for (int i = tid; i < V; i += kSampThreads)
      if (order_key(val[i]) < kth) val[i] = -INFINITY;
    __syncthreads();
  }

  // ---- 3. top-p (nucleus) on the sorted distribution
  if (a.p.top_p < 1.0f) {
    int Vp = 1;
    while (Vp < V) Vp <<= 1;
    for (int i = tid; i < Vp; i += kSampThreads)
      keys[i] = i < V ? ((unsigned long long)order_key(val[i]) << 32) | (unsigned)(~(unsigned)i) : 0ull;  // desc value, asc index
    __syncthreads();
    for (int size = 2; size <= Vp; size <<= 1)
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (Vp >> 1); t += kSampThreads) {
          int lo = 2 * t - (t & (stride - 1));
          int hi = lo + stride;
          bool desc = (lo & size) == 0;
          unsigned long long x = keys[lo], y = keys[hi];
          if ((x < y) == desc) { keys[lo] = y; keys[hi] = x; }
        }
        __syncthreads();
      }
    // sorted descending.  softmax over sorted values, prefix sums in sorted order.
    const int per = (Vp + kSampThreads - 1) / kSampThreads;
    const int i0 = tid * per;
    float local = 0.f;
    for (int j = 0; j < per; ++j) {
      int i = i0 + j;
      if (i < V) { int id = (int)~(unsigned)(keys[i] & 0xffffffffu); local += expf(val[id] - mx); }
    }
    float total = red.sum(local);
    // exclusive prefix of `local` across threads
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();
    if (lane == 31) s_redf[warp] = incl;
    __syncthreads();
    float wbase = 0.f;
    for (int w = 0; w < warp; ++w) wbase += s_redf[w];
    float run = wbase + incl - local;   // mass strictly before this thread's first entry
    if (tid == 0) s_cut = keys[0];
    __syncthreads();
    // keep entry i iff i == 0 or cumsum[i-1] <= top_p; kept entries form a prefix -> remember the last kept key
    unsigned long long my_cut = ~0ull;
    for (int j = 0; j < per; ++j) {
      int i = i0 + j;
      if (i < V) {
        bool keep = (i == 0) || (run / total <= a.p.top_p);
        int id = (int)~(unsigned)(keys[i] & 0xffffffffu);
        if (keep) my_cut = keys[i];
        run += expf(val[id] - mx);
      }
    }
    atomicMin(&s_cut, my_cut);
    __syncthreads();
    const unsigned long long cut = s_cut;
    __syncthreads();
    for (int i = tid; i < V; i += kSampThreads) {
      unsigned long long kk = ((unsigned long long)order_key(val[i]) << 32) | (unsigned)(~(unsigned)i);
      if (kk < cut) val[i] = -INFINITY;
    }
    __syncthreads();
  }

  if (a.filtered_out)
    for (int i = tid; i < V; i += kSampThreads) a.filtered_out[(size_t)b * V + i] = val[i];

  // ---- 4. softmax denominator (only when probabilities are part of the result)
  const bool draw = a.p.sample_logits != 0;
  const bool exact = a.probs_out != nullptr || a.noise != nullptr || !draw;
  float denom = 1.f;
  if (a.probs_out) {
    float lsum = 0.f;
    for (int i = tid; i < V; i += kSampThreads) lsum += expf(val[i] - mx);
    denom = red.sum(lsum);
  }

  // ---- 5. pick
  const unsigned long long seed = a.seed_ptr ? *a.seed_ptr : (unsigned long long)a.p.seed;
  const uint2 pkey = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint64_t off = a.p.offset + (uint64_t)step;
  float best = -INFINITY;
  int besti = 0x7fffffff;
  if (exact) {
    // API path: argmax(probs) or argmax(probs / noise) with IEEE divisions, as torch evaluates it
    for (int i = tid; i < V; i += kSampThreads) {
      const float p = __fdiv_rn(expf(val[i] - mx), denom);
      if (a.probs_out) a.probs_out[(size_t)b * V + i] = p;
      float score = p;
      if (draw) {
        float e = a.noise ? a.noise[(size_t)b * V + i] : 1.f;
        if (!a.noise) {
          const uint4 r4 = philox4x32_10(make_uint4((uint32_t)(i >> 2), (uint32_t)b, (uint32_t)off, (uint32_t)(off >> 32)), pkey);
          const uint32_t rb[4] = {r4.x, r4.y, r4.z, r4.w};
          e = exp1_from_bits(rb[i & 3]);
        }
        score = __fdiv_rn(p, e);
      }
      if (score > best || (score == best && i < besti)) { best = score; besti = i; }
    }
  } else {
    // decode-loop path: one multinomial draw = argmax_i p_i / E_i, E_i ~ Exp(1) from Philox.  The softmax denominator
    // is a common factor and the winner is decided in the log domain: (x_i - max) - log(E_i).
    for (int i4 = tid * 4; i4 < V; i4 += kSampThreads * 4) {
      const uint4 r4 = philox4x32_10(make_uint4((uint32_t)(i4 >> 2), (uint32_t)b, (uint32_t)off, (uint32_t)(off >> 32)), pkey);
      const uint32_t rb[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i4 + j;
        if (i >= V) break;
        const float x = val[i];
        if (x == -INFINITY) continue;                                  // filtered out
        const float score = (x - mx) - __logf(exp1_from_bits(rb[j]));
        if (score > best) { best = score; besti = i; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  __syncthreads();
  if (lane == 0) { s_redf[warp] = best; s_redi[warp] = besti; }
  __syncthreads();
  if (warp == 0) {
    best = s_redf[lane];
    besti = s_redi[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
      if (besti == 0x7fffffff) besti = 0;
      if (a.idx_out) a.idx_out[b] = besti;
      if (a.seq_out) a.seq_out[(size_t)b * a.seq_stride + step] = besti;
      // last CTA to finish advances the device-side loop state (every CTA has read it by now)
      if (a.done_ctr) {
        __threadfence();
        unsigned t = atomicAdd(a.done_ctr, 1u);
        if (t == gridDim.x - 1) {
          *a.done_ctr = 0;
          if (a.step_ptr) *a.step_ptr = step + 1;
          if (a.pos_ptr) *a.pos_ptr = *a.pos_ptr + 1;
        }
      }
    }
  }
}

int launch_sample(const SampleArgs& a, cudaStream_t st) {
  VLG_REQUIRE(a.logits && a.B > 0 && a.V > 0, "sample: bad arguments");
  VLG_REQUIRE(a.rows == a.B || a.rows == 2 * a.B, "sample: rows (%d) must be B or 2*B (B=%d)", a.rows, a.B);
  size_t smem = (((size_t)a.V * 4 + 15) & ~size_t(15));
  if (a.p.top_p < 1.0f) {
    size_t Vp = 1;
    while (Vp < (size_t)a.V) Vp <<= 1;
    smem += Vp * 8;
  }
  VLG_REQUIRE(smem <= 220 * 1024, "sample: vocabulary %d too large for the shared-memory sampler%s", a.V,
              a.p.top_p < 1.0f ? " with top-p" : "");
  static size_t attr[kMaxDevices] = {};
  const int dev = current_device();
  if (smem > attr[dev]) {
    VLG_CUDA(cudaFuncSetAttribute(cfg_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dev] = smem;
  }
  VLG_CUDA(launch_kernel(cfg_sample_kernel, dim3(a.B), dim3(kSampThreads), smem, st, a));
  VLG_LAUNCH_CHECK();
  return VLG_OK;
}

}  // namespace vlg

extern "C" int vlg_cfg_sample(const float* logits_dev, int rows, int B, int V, const vlg_sample_params* p,
                              const float* noise_dev, int64_t* idx_dev, float* probs_dev, void* stream) {
  using namespace vlg;
  VLG_REQUIRE(logits_dev && p && idx_dev, "vlg_cfg_sample: null pointer");
  SampleArgs a{};
  a.logits = logits_dev;
  a.rows = rows;
  a.B = B;
  a.V = V;
  a.p = *p;
  a.noise = noise_dev;
  a.idx_out = idx_dev;
  a.probs_out = probs_dev;
  a.cfg_interval = -1;
  return launch_sample(a, (cudaStream_t)stream);
}

extern "C" int vlg_top_k_top_p(const float* logits_dev, float* out_dev, int B, int V, int top_k, float top_p,
                               void* stream) {
  using namespace vlg;
  VLG_REQUIRE(logits_dev && out_dev, "vlg_top_k_top_p: null pointer");
  SampleArgs a{};
  a.logits = logits_dev;
  a.rows = B;
  a.B = B;
  a.V = V;
  a.p.cfg_scale = 1.f;
  a.p.temperature = 1.f;
  a.p.top_k = top_k;
  a.p.top_p = top_p;
  a.filtered_out = out_dev;
  a.cfg_interval = -1;
  return launch_sample(a, (cudaStream_t)stream);
}
