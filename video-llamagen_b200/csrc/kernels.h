// Internal launcher declarations shared between the .cu translation units.
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace vlg {

int launch_rmsnorm(int dtype, const void* x, const void* w, void* out, int rows, int dim, float eps, cudaStream_t st);

int launch_gather(int dtype, const void* table, void* out, int rows, int dim, const int64_t* idx64,
                  const int32_t* seq, int seq_stride, const int* col_ptr, int col_off, const int64_t* labels, int B,
                  int null_class, int table_rows, cudaStream_t st);

int launch_gemm(int dtype, int epi, const void* x, const void* w, int rows, int N, int K, const EpiParams& P,
                cudaStream_t st);

// Single-position attention over the cache prefix.  pos comes from *pos_ptr when non-null.
struct AttnArgs {
  const void* q;        // [rows, H, D]
  const void* k_cache;  // [rows, H, S_max, D]
  const void* v_cache;
  void* out;            // [rows, H*D]
  int rows, H, D, S_max;
  const int* pos_ptr;
  int pos_imm;
  const uint8_t* prefix_mask;  // [rows, T_prefix] or null
  int T_prefix;
  float* scratch;       // split partials
  size_t scratch_bytes;
  int max_pos;          // upper bound of pos over the life of a captured graph (sizes the split count)
  bool early_prefetch;  // TMA kernel only: the producer may read the position and stream old cache rows before
                        // griddepcontrol.wait.  Safe when everything older than the immediate predecessor has completed by
                        // the time this kernel's CTAs run -- true in the decode step, where the predecessor (layer-tail or
                        // QKV GEMM kernel, ~200 KB of shared memory per CTA) cannot share an SM with this kernel, so its CTAs
                        // have left their own wait long before; false for callers with an unknown predecessor
  int q_per_row;        // generic kernel only, batched prefill: q / out hold q_per_row queries per cache row (positions
                        // pos_imm .. pos_imm + q_per_row - 1); 0 = one query per cache row
};
int launch_attn_generic(int dtype, const AttnArgs& a, cudaStream_t st);
size_t attn_scratch_bytes(int rows, int H, int D, int S_max);
// bf16 only: persistent, TMA-fed streaming kernel (attn_tma.cu)
int launch_attn_tma(int dtype, const AttnArgs& a, cudaStream_t st);

// bf16, rows <= 128: everything between two attentions as one persistent kernel (layer_tail.cu).
//   wo != null  : attn -> wo + residual -> RMSNorm(norm1_w) -> w13 -> SwiGLU -> w2 + residual, then RMSNorm(norm2_w) -> xn
//   wo == null  : RMSNorm(norm2_w) of h -> xn only (the launch in front of layer 0)
//   wqkv != null: xn -> wqkv -> RoPE + q / KV-cache write described by `rope`
struct LtLaunch {
  int rows, dim, hidden;
  const void* attn;     // [rows, dim]
  void* h;              // [rows, dim] residual stream, in place
  void* xn;             // [rows, dim]
  void* ffn;            // [rows, hidden]
  const void *wo, *w13, *w2, *wqkv;
  const void *norm1_w, *norm2_w;
  float eps;
  EpiParams rope;
  float* part;          // split-K scratch, layer_tail_scratch_bytes()
  size_t part_bytes;
  unsigned* bar;        // 2 counters, zero
  unsigned* err = nullptr;   // device-visible pointer to a mapped host word: grid-barrier timeout flag (may be null)
  // optional L2 prefetch for the kernel that follows (layer_tail.cu, lt_pf_slice): next layer's K/V prefix rows [0, pos)
  const void *pf_k = nullptr, *pf_v = nullptr;
  int pf_n_rh = 0;
  long long pf_stride_bytes = 0;
  const int* pf_pos_ptr = nullptr;
  int pf_pos_imm = 0;
  const void* pf_flat = nullptr;   // and / or one flat region (head weights)
  size_t pf_flat_bytes = 0;
};
bool layer_tail_supported(int rows, int dim, int hidden);
size_t layer_tail_scratch_bytes(int max_rows, int dim, int hidden);
int launch_layer_tail(const LtLaunch& a, cudaStream_t st);
// Fused CFG mix + temperature + top-k/top-p + softmax + pick.
struct SampleArgs {
  const float* logits;  // [rows, V]
  int rows, B, V;
  vlg_sample_params p;
  const float* noise;   // [B, V] or null
  int64_t* idx_out;     // [B] or null
  float* probs_out;     // [B, V] or null
  float* filtered_out;  // [B, V] or null: logits after temperature and top-k/top-p (-inf where removed)
  // decode-loop extras (all optional)
  int32_t* seq_out;     // [B, seq_stride]: token written at column *step_ptr
  int seq_stride;
  int* step_ptr;        // device step counter: read for the column / RNG offset / cfg_interval, then incremented
  int* pos_ptr;         // device cache position: incremented together with the step
  int cfg_interval;     // -1: always guided
  unsigned* done_ctr;   // block-completion counter for the increment
  const unsigned long long* seed_ptr;   // device-resident Philox seed (overrides p.seed when non-null)
};
int launch_sample(const SampleArgs& a, cudaStream_t st);

}  // namespace vlg
