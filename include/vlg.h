/* vlg.h -- C ABI of libvlg.so: the B200 (sm_100a) decode hot path of Video-LlamaGen.
 *
 * The reference (BinZhu-ece/Video-LlamaGen) is pure Python/PyTorch and has no FFI of
 * its own; the boundary it offers is the Python surface of
 *   autoregressive/models/gpt.py       (Transformer.setup_caches / forward, :318-385)
 *   autoregressive/models/generate.py  (generate / prefill / decode_* / sample, :16-180)
 *   tokenizer/tokenizer_image/vq_model.py (VectorQuantizer.forward / get_codebook_entry, :215-276)
 * Each entry point below names the reference lines it replaces; INTEGRATION.md shows the
 * ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; every pointer named *_dev is device memory owned by the caller
 *     (a torch tensor kept alive by the Python side); *_host is host memory.
 *   - every call returns VLG_OK or a negative vlg_status; vlg_last_error() gives the text.
 *   - no call allocates or frees device memory except vlg_ctx_create/destroy (which own
 *     only small descriptor/graph objects; the KV cache and activations live in the
 *     caller-provided workspace) -- and no call synchronises the device unless its
 *     comment says so.  `stream` is a cudaStream_t passed as void*.
 *   - one host thread per context, one process per GPU (generate() in the reference is not
 *     re-entrant either: it mutates model.causal_mask / KV caches, generate.py:154-165).
 */
#ifndef VLG_H_
#define VLG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VLG_ABI_VERSION 1

typedef enum {
  VLG_OK = 0,
  VLG_ERR_ARG = -1,         /* bad argument / unsupported shape */
  VLG_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed */
  VLG_ERR_UNSUPPORTED = -3, /* device is not sm_100 */
  VLG_ERR_STATE = -4        /* call sequence violated (e.g. decode before prefill) */
} vlg_status;

typedef enum { VLG_F32 = 0, VLG_BF16 = 1 } vlg_dtype;
typedef enum { VLG_C2I = 0, VLG_T2I = 1, VLG_T2V = 2 } vlg_model_type;

int vlg_abi_version(void);
const char* vlg_last_error(void);

/* ------------------------------------------------------------------------------------------
 * VQ tokenizer index search -- replaces VectorQuantizer.forward (vq_model.py:215-259, eval mode)
 * and get_codebook_entry (vq_model.py:261-276).
 * ------------------------------------------------------------------------------------------ */

/* z_dev        fp32 [B, C, HW]  (the [B,C,H,W] activations the reference passes in, C = e_dim)
 * codebook_dev fp32 [n_e, C]    raw embedding.weight (normalised inside when l2_norm != 0)
 * idx_dev      int64 [B*HW]     argmin_j ( |z|^2 + |e_j|^2 - 2 z.e_j ), first index on ties
 * zq_dev       fp32 [B, C, HW]  selected (normalised) codebook rows, channel-first; may be NULL
 * scratch_dev  fp32 [vlg_vq_scratch_floats(B*HW, C, n_e)] work area, 16-byte aligned: normalised codebook, |e_j|^2, the
 *              duplicated / padded codebook layout the score kernel streams, and the list of latents whose two best codes
 *              are closer than fp32 rounding (settled by an exact second pass) */
size_t vlg_vq_scratch_floats(int64_t N, int C, int n_e);
int vlg_vq_argmin(const float* z_dev, const float* codebook_dev, int64_t* idx_dev, float* zq_dev,
                  float* scratch_dev, int64_t B, int C, int64_t HW, int n_e, int l2_norm, void* stream);

/* out_dev fp32: channel_first ? [B, C, HW] : [B*HW, C]  = normalise(codebook)[idx] */
int vlg_codebook_gather(const int64_t* idx_dev, const float* codebook_dev, float* out_dev, int64_t B, int C,
                        int64_t HW, int n_e, int l2_norm, int channel_first, void* stream);

/* First step of VQModel.decode_code (vq_model.py:53-55 -> :47-48): get_codebook_entry fused with the 1x1
 * post_quant_conv, so the [B, C, HW] quantised map is never written.
 * out_dev fp32 [B, O, HW] = bias[o] + sum_c weight[o, c] * normalise(codebook)[idx[b*HW + hw], c]
 * weight_dev fp32 [O, C] (post_quant_conv.weight viewed 2-D), bias_dev fp32 [O] or NULL.          */
int vlg_codebook_project(const int64_t* idx_dev, const float* codebook_dev, const float* weight_dev,
                         const float* bias_dev, float* out_dev, int64_t B, int C, int64_t HW, int n_e, int O,
                         int l2_norm, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sampling -- replaces decode_one_token's CFG mix (generate.py:95-99), top_k_top_p_filtering
 * (generate.py:16-54) and sample (generate.py:57-66).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  float cfg_scale;     /* logits = uncond + (cond - uncond) * cfg_scale when rows == 2*B and guided */
  int guided;          /* 0: take the conditional half unchanged (cfg_interval passed, generate.py:98-99) */
  float temperature;   /* logits / max(temperature, 1e-5) */
  int top_k;           /* 0 = off; threshold = k-th largest, strict '<' removed (ties survive) */
  float top_p;         /* 1.0 = off */
  int sample_logits;   /* 1: multinomial (argmax of probs / Exp(1));  0: argmax(probs) */
  uint64_t seed;       /* Philox seed for the in-kernel Exp(1) draws (used when noise_dev == NULL) */
  uint64_t offset;     /* per-call stream offset (the decode loop passes the step index) */
} vlg_sample_params;

/* logits_dev fp32 [rows, V]   rows = 2*B (cond rows first, then uncond) or rows = B
 * noise_dev  fp32 [B, V] Exp(1) draws, or NULL to draw in-kernel (Philox4x32-10)
 * idx_dev    int64 [B]        chosen token
 * probs_dev  fp32 [B, V] softmax after filtering, or NULL to skip the write                   */
int vlg_cfg_sample(const float* logits_dev, int rows, int B, int V, const vlg_sample_params* p,
                   const float* noise_dev, int64_t* idx_dev, float* probs_dev, void* stream);

/* top_k_top_p_filtering alone (generate.py:16-54): out_dev fp32 [B, V] = logits with removed entries at -inf */
int vlg_top_k_top_p(const float* logits_dev, float* out_dev, int B, int V, int top_k, float top_p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decode context -- replaces Transformer.setup_caches (gpt.py:318-332), the inference branch of
 * Transformer.forward (gpt.py:349-371; gpt_video.py:403-431 for t2v) and the generate() loop
 * (generate.py:127-180).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int dim, n_layer, n_head, hidden_dim;
  int vocab_size;      /* V (logit width); unused for t2v */
  int cls_token_num;   /* T: 1 for c2i, 120 for t2i/t2v */
  int num_classes;     /* c2i: the null class id used for the unconditional rows (generate.py:131) */
  int caption_dim;     /* t2i/t2v condition width (2048) */
  int vae_embed_dim;   /* t2v latent width (8) */
  int model_type;      /* vlg_model_type */
  int dtype;           /* vlg_dtype of weights, activations and KV cache */
  float norm_eps;
} vlg_geometry;

typedef struct {
  const void* attention_norm; /* [dim] */
  const void* wqkv;           /* [3*dim, dim]   attention.wqkv.weight */
  const void* wo;             /* [dim, dim] */
  const void* ffn_norm;       /* [dim] */
  const void* w13;            /* [2*hidden, dim] rows interleaved: 2j = w1[j], 2j+1 = w3[j] */
  const void* w2;             /* [dim, hidden] */
} vlg_layer_weights;

typedef struct {
  const void* tok_embeddings;  /* [V, dim]                         (c2i, t2i) */
  const void* cls_table;       /* [num_classes+1, dim]             (c2i) */
  const void* cap_fc1;         /* [dim, caption_dim]               (t2i, t2v) */
  const void* cap_fc2;         /* [dim, dim] */
  const void* adapter_in_fc1;  /* [dim, vae_embed_dim]             (t2v) vae_latent_adapter */
  const void* adapter_in_fc2;  /* [dim, dim] */
  const void* adapter_out_fc1; /* [dim, dim]                       (t2v) vae_latent_adapter2 */
  const void* adapter_out_fc2; /* [vae_embed_dim, dim] */
  const void* norm;            /* [dim] */
  const void* output;          /* [V, dim] */
  const float* rope_table;     /* fp32 [rope_rows, head_dim/2, 2] (cos, sin): gpt.py:407-420 / gpt_video.py:532-552 */
  int rope_rows;
  const vlg_layer_weights* layers; /* [n_layer] (host array, copied) */
} vlg_weights;

typedef struct vlg_ctx vlg_ctx;

/* bytes of device workspace a context needs for `max_rows` rows (2*B with CFG) and `max_seq` slots */
size_t vlg_ctx_workspace_bytes(const vlg_geometry* g, int max_rows, int max_seq);

/* Weight pointers are borrowed (device memory, dtype g->dtype, layouts above) and must outlive the
 * context.  workspace_dev: >= vlg_ctx_workspace_bytes, 1024-byte aligned.  Synchronises once. */
int vlg_ctx_create(const vlg_geometry* g, const vlg_weights* w, int max_rows, int max_seq, void* workspace_dev,
                   size_t workspace_bytes, vlg_ctx** out);
void vlg_ctx_destroy(vlg_ctx* ctx);

/* Start a sequence batch: `rows` cache rows (cond rows then uncond rows when CFG is on).
 * prefix_mask_dev: uint8 [rows, cls_token_num], 1 = key visible (emb_masks, generate.py:156-165), or NULL. */
int vlg_begin(vlg_ctx* ctx, int rows, const uint8_t* prefix_mask_dev, void* stream);

/* One position of Transformer.forward's inference branch for all rows (seqlen 1):
 * exactly one of the three inputs is non-NULL
 *   tokens_dev  int64 [rows]                      -> tok_embeddings gather   (gpt.py:354)
 *   labels_dev  int64 [rows]                      -> cls_embedding gather    (gpt.py:352, c2i)
 *   embeds_dev  dtype [rows, dim]                 -> already-projected prefix embedding / adapter output
 * pos: cache slot / RoPE row.  logits_dev: fp32 [rows, V] (or NULL to skip the head).
 * hidden_dev: dtype [rows, dim] final-normed hidden state (or NULL).                          */
int vlg_forward_pos(vlg_ctx* ctx, const int64_t* tokens_dev, const int64_t* labels_dev, const void* embeds_dev,
                    int pos, float* logits_dev, void* hidden_dev, void* stream);

/* The whole conditioning prefix in one call (t2i / t2v: gpt.py:351-352 at S = cls_token_num positions, the reference's
 * prefill shape, generate.py:77-86): embeds_dev dtype [rows, T, dim] = the caption projection of every row (vlg_mlp which = 0),
 * positions 0..T-1, causal attention inside the prefix under the mask given to vlg_begin.  K/V of the prefix land in the
 * caches; hidden_dev dtype [rows, dim] / logits_dev fp32 [rows, V] receive the LAST position's final-normed hidden state /
 * logits (either may be NULL) -- the only position the reference's callers use.  Real GEMMs over rows*T rows instead of T
 * single-position passes. */
int vlg_forward_prefix(vlg_ctx* ctx, const void* embeds_dev, float* logits_dev, void* hidden_dev, void* stream);

/* Caption / latent projection MLP fc2(gelu_tanh(fc1 x)) (gpt.py:118-131): which = 0 cap_proj,
 * 1 vae_latent_adapter, 2 vae_latent_adapter2.  x_dev dtype [n, in], out_dev dtype [n, out].
 * bf16 only: `in` is rounded up to a multiple of 64 and `out` to a multiple of 16 (tensor-core tile granularity): the
 * caller zero-pads x / fc1 columns and fc2 rows / out columns accordingly (only the 8-wide t2v adapter ends are affected). */
int vlg_mlp(vlg_ctx* ctx, int which, const void* x_dev, void* out_dev, int n, void* stream);

/* The whole of generate() for c2i / t2i with device-resident inputs (generate.py:127-180):
 *   c2i: cond_dev int64 [B] class labels;  t2i: cond_dev dtype [B, T, caption_dim] (+ uncond rows built inside
 *   from uncond_dev dtype [T, caption_dim] when cfg_scale > 1) and emb_masks_dev uint8 [B, T] or NULL.
 * tokens_dev int32 [B, max_new_tokens] receives the sequence.  No host synchronisation. */
typedef struct {
  int B;
  int max_new_tokens;
  float cfg_scale;     /* > 1 doubles the rows (generate.py:130-141) */
  int cfg_interval;    /* -1 = guide every step (generate.py:113-114) */
  float temperature;
  int top_k;
  float top_p;
  int sample_logits;
  uint64_t seed;
} vlg_generate_params;

int vlg_generate(vlg_ctx* ctx, const vlg_generate_params* p, const void* cond_dev, const void* uncond_dev,
                 const uint8_t* emb_masks_dev, int32_t* tokens_dev, void* stream);

/* t2v (gpt_video.py trunk, SURVEY.md 8a row a17 / BASELINE cfg-4): the loop the reference's samplers call as
 * `generate(gpt_model, c_indices, vae_t * latent_size**2, c_emb_masks, cfg_scale=...)` (sample_t2v_1f.py:199-205; the
 * module it imports, models/generate_video.py, is missing from the fork).  The model emits a continuous latent per
 * position -- no logits, no sampler: prefill over the T caption positions, then N-1 steps of
 *   latent_s = cfg_mix(vae_latent_adapter2(hidden))  ->  next input = vae_latent_adapter(latent_s).
 * cond_dev dtype [B, T, caption_dim]; uncond_dev dtype [T, caption_dim] (needed when cfg_scale > 1); emb_masks_dev
 * uint8 [B, T] or NULL; latents_dev dtype [B, max_new_tokens, vae_embed_dim].  temperature / top_k / top_p /
 * sample_logits / seed of the params are ignored.  bf16: vae_latent_adapter.fc1 zero-padded to 64 input columns and
 * vae_latent_adapter2.fc2 to 16 output rows by the caller (tensor-core tile granularity). */
int vlg_generate_t2v(vlg_ctx* ctx, const vlg_generate_params* p, const void* cond_dev, const void* uncond_dev,
                     const uint8_t* emb_masks_dev, void* latents_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * Diffusion-head token sampler -- replaces DiffLoss.sample (autoregressive/models/diffloss.py:35-52):
 * the `steps`-long p_sample_loop (diffusion/gaussian_diffusion.py:376-517, respaced by diffusion/respace.py:63-129) over
 * SimpleMLPAdaLN (diffloss.py:151-247), eps-prediction with LEARNED_RANGE variance (diffusion/__init__.py:11-47), fp32.
 * Call site in the reference: generate_video_diff.py:88-90 and :130-132 (one call per generated token).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int C;      /* target_channels: width of the latent being sampled (vae_embed_dim, 8 for the reference's VAE); <= 16 */
  int W;      /* model_channels (diffloss_w, 1024): 128, 256, 512, 768 or 1024 */
  int Z;      /* z_channels: width of the conditioning vector (the trunk's dim) */
  int depth;  /* residual blocks (diffloss_d, 3) */
  int steps;  /* respaced sampling steps (num_sampling_steps, 100) */
} vlg_diff_geometry;

/* all fp32 device pointers in nn.Linear layout [out, in]; names = state_dict keys under `net.` */
typedef struct {
  const float *ln_w, *ln_b;   /* res_blocks.b.in_ln.{weight,bias}            [W]       */
  const float *w1, *b1;       /* res_blocks.b.mlp.0.{weight,bias}            [W, W], [W] */
  const float *w2, *b2;       /* res_blocks.b.mlp.2.{weight,bias}            [W, W], [W] */
  const float *ada_w, *ada_b; /* res_blocks.b.adaLN_modulation.1.{weight,bias} [3W, W], [3W] */
} vlg_diff_block;

typedef struct {
  const float *time_w0, *time_b0;         /* time_embed.mlp.0  [W, 256], [W] */
  const float *time_w1, *time_b1;         /* time_embed.mlp.2  [W, W], [W] */
  const float *cond_w, *cond_b;           /* cond_embed        [W, Z], [W] */
  const float *in_w, *in_b;               /* input_proj        [W, C], [W] */
  const vlg_diff_block* blocks;           /* [depth] (host array, copied) */
  const float *final_ada_w, *final_ada_b; /* final_layer.adaLN_modulation.1 [2W, W], [2W] */
  const float *final_w, *final_b;         /* final_layer.linear [2C, W], [2C] */
} vlg_diff_weights;

typedef struct vlg_diff vlg_diff;

size_t vlg_diff_workspace_bytes(const vlg_diff_geometry* g, int max_rows);   /* 0 when the geometry is unsupported */

/* timestep_map_host int32 [steps]: original timestep of respaced step i (respace.py:77-85).
 * tables_host fp32 [6][steps], the float64 tables of gaussian_diffusion.py:153-201 rounded to fp32 as _extract_into_tensor
 * does (:865-877): sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
 * posterior_log_variance_clipped, log(betas).  Both are schedule constants the caller derives on the host, as the
 * reference does in numpy.  Weight pointers are borrowed and must outlive the handle; workspace_dev is 256-byte aligned.
 * Computes the model's timestep-embedding table on the device and synchronises once.                              */
int vlg_diff_create(const vlg_diff_geometry* g, const vlg_diff_weights* w, const int32_t* timestep_map_host,
                    const float* tables_host, int max_rows, void* workspace_dev, size_t workspace_bytes, vlg_diff** out);
void vlg_diff_destroy(vlg_diff* d);

/* One DiffLoss.sample call: z_dev fp32 [rows, Z] -> out_dev fp32 [rows, C].
 * cfg_scale == 1: plain sampling.  cfg_scale != 1: forward_with_cfg (diffloss.py:237-247) -- rows [0, rows/2) are the
 * conditional, [rows/2, rows) the unconditional copies; both halves are fed the first half's sample, eps is mixed
 * uncond + cfg_scale * (cond - uncond), every row keeps its own variance output and its own noise.
 * noise_dev fp32 [rows, C]: the start sample (the caller duplicates it across the halves under guidance, as
 * diffloss.py:38-39 does), or NULL to draw it in-kernel (Philox4x32-10 keyed by seed/offset, duplicated likewise).
 * step_noise_dev fp32 [steps, rows, C]: entry i is the draw added at loop index i (i runs steps-1 .. 0; index 0 is
 * never used, gaussian_diffusion.py:413-415), or NULL to draw in-kernel.  No host synchronisation.                */
int vlg_diff_sample(vlg_diff* d, const float* z_dev, int rows, float temperature, float cfg_scale, const float* noise_dev,
                    const float* step_noise_dev, uint64_t seed, uint64_t offset, float* out_dev, void* stream);

/* Synchronises `stream` and reports (VLG_ERR_STATE) whether a launch since the last check gave up in a grid barrier. */
int vlg_diff_check(vlg_diff* d, void* stream);

/* Measurement only: register-resident FP32 FMA loop on every SM (variant 0: FFMA, 1: packed FFMA2); the caller times
 * it with events.  *flops_out = floating-point operations of the launch.  The roofline denominator of vlg_vq_argmin. */
int vlg_debug_fp32_fma_peak(int variant, int iters, float* out_dev, long long* flops_out, void* stream);

/* Synchronises `stream` and reports whether any layer-tail launch of this context since the last check gave up in its
 * grid barrier (VLG_ERR_STATE: its CTAs were not co-resident because other work occupied SMs of the GPU; the outputs are
 * invalid and the call should be repeated once the GPU is free).  Every other context call performs the same check,
 * without synchronising, on entry; vlg_generate_c2i_host also on exit. */
int vlg_ctx_check(vlg_ctx* ctx, void* stream);

/* Same with HOST buffers (pinned or pageable): copies cond in, runs, copies tokens out and synchronises
 * the stream.  This is the call the end-to-end benchmark times. c2i only (labels int64 [B]). */
int vlg_generate_c2i_host(vlg_ctx* ctx, const vlg_generate_params* p, const int64_t* labels_host,
                          int32_t* tokens_host, void* stream);

/* number of kernels launched by this library since load (for bench.py's gpu_launches) */
uint64_t vlg_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Building blocks exposed for unit parity tests (each mirrors one reference op).
 * dtype = vlg_dtype of x / w / out unless stated.
 * ------------------------------------------------------------------------------------------ */
/* RMSNorm (gpt.py:143-148): out = T( T(x * rsqrt(mean(x^2)+eps)) * w ) */
int vlg_rmsnorm(const void* x_dev, const void* w_dev, void* out_dev, int rows, int dim, float eps, int dtype,
                void* stream);

/* out[rows,N] = x[rows,K] . w[N,K]^T  (nn.Linear, bias-free).  bf16: tcgen05 tensor cores, fp32 accumulate,
 * one rounding to bf16.  fp32: CUDA-core FMA, no TF32.  out_f32 != 0 widens the bf16 result to fp32
 * (`.float()` after the head, gpt.py:371). */
int vlg_linear(const void* x_dev, const void* w_dev, void* out_dev, int rows, int N, int K, int dtype, int out_f32,
               void* stream);

/* Single-position attention over the cache prefix [0, pos] (gpt.py:221-239 at seqlen 1, math SDPA):
 * q_dev dtype [rows, H, D]; k_cache/v_cache dtype [rows, H, S_max, D]; prefix_mask as in vlg_begin;
 * out_dev dtype [rows, H*D].  impl: 0 = auto, 1 = generic (any dtype), 2 = TMA streaming (bf16 only). */
int vlg_attn_decode(const void* q_dev, const void* k_cache_dev, const void* v_cache_dev, void* out_dev, int rows,
                    int H, int D, int S_max, int pos, const uint8_t* prefix_mask_dev, int T_prefix, int dtype,
                    int impl, void* scratch_dev, size_t scratch_bytes, void* stream);
size_t vlg_attn_scratch_bytes(int rows, int H, int D, int S_max);

/* instrumentation (VLG_GEMM_DEBUG & 4): globaltimer phase stamps of CTA 0 of the last tcgen05 GEMM; synchronises */
int vlg_debug_gemm_timestamps(unsigned long long* out16_host);
/* measurement only: the layer-tail launches of one position (front + one per layer), attention skipped; bf16, rows <= 128 */
int vlg_debug_trunk_without_attention(vlg_ctx* ctx, int pos, void* stream);
/* instrumentation (VLG_LT_DEBUG=1): globaltimer phase stamps [160 CTAs][32 slots] of the last full layer-tail launch */
int vlg_debug_lt_timestamps(unsigned long long* out_host);
/* verification of the CTA-pair (tcgen05 cta_group::2) GEMM building block: out [rows, N] fp32 = x [rows, K] . w [N, K]^T, bf16
 * operands; rows <= 128, K % 64 == 0, bnp in {64, 128} the pair's column tile, N % bnp == 0 */
int vlg_debug_pair_gemm(const void* x_dev, const void* w_dev, float* out_dev, int rows, int N, int K, int bnp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLG_H_ */
