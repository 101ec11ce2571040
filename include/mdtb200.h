/*
 * mdtb200.h -- C ABI of the B200-native MDT denoising hot path (libmdtb200.so).
 *
 * Drop-in boundary: everything below replaces the arithmetic behind the reference's
 * Hydra-built score wrapper (paths relative to the reference repo root):
 *
 *   mdt/models/edm_diffusion/score_wrappers.py:18-100   GCDenoiser.{forward,loss,forward_context_only}
 *   mdt/models/networks/mdtv_transformer.py:208-244     MDTVTransformer.{forward,forward_enc_only,forward_dec_only}
 *   mdt/models/edm_diffusion/gc_sampling.py:164-210,256-311,699-733,922-951   sample_{euler,heun,dpmpp_2m,ddim}
 *   mdt/models/mdtv_agent.py:523-550,593-658            denoise_actions / sample_loop (the caller)
 *
 * Conventions
 *   - plain C types only: device/host pointers, sizes, a cudaStream_t passed as void*.
 *   - all tensors fp32, contiguous, row-major; weights in nn.Linear layout (out, in).
 *   - every function returns 0 on success, a negative MDTB200_E* code on failure and never
 *     throws / exits; mdtb200_last_error() gives the message.  Asynchronous CUDA faults
 *     surface at the caller's next synchronisation, as in PyTorch.
 *   - input/output pointers are borrowed for the duration of the call only.  Weight pointers
 *     are read by mdtb200_commit_weights() and not retained: the library keeps its own packed
 *     copy (re-commit after load_state_dict / EMA swap / optimizer step).
 *   - a handle is bound to the CUDA device that was current at mdtb200_create() and holds no
 *     global mutable state.  It owns ONE static workspace: calls on the same handle must be
 *     ordered (one host thread at a time, and stream-ordered on the device -- issue them on one
 *     stream, or synchronise between streams); use one handle per concurrent stream / thread.
 *     All work is launched on the stream passed in (sampling graphs are captured on a private
 *     stream and *launched* on the caller's stream, so the legacy default stream is fine).
 *     The graph cache is bounded (16 entries, least recently used evicted).
 */
#ifndef MDTB200_H_
#define MDTB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDTB200_ABI_VERSION 1

#if defined(__GNUC__)
#define MDTB200_API __attribute__((visibility("default")))
#else
#define MDTB200_API
#endif

enum {
  MDTB200_OK = 0,
  MDTB200_EINVAL = -1,     /* bad argument / shape / config */
  MDTB200_ESTATE = -2,     /* call order violated (weights not committed, no context encoded, ...) */
  MDTB200_ECUDA = -3,      /* CUDA runtime / driver error (message carries cudaGetErrorString) */
  MDTB200_ENOMEM = -4,
  MDTB200_EUNSUPPORTED = -5
};

/* score-network flavour: conf/model/model/mdtv_transformer.yaml vs mdt_transformer.yaml */
enum { MDTB200_VARIANT_MDTV = 0, MDTB200_VARIANT_MDT = 1 };

/* GEMM arithmetic.  FP32 = exact fp32 FMA on CUDA cores (reference-grade rounding).
 * BF16X3 = tcgen05 tensor cores, each fp32 operand split into bf16 hi+lo, three MMAs
 * (hi*hi + lo*hi + hi*lo) accumulated in fp32 TMEM: ~2^-17 relative operand error, meets
 * the 1e-4 action tolerance.  BF16 = single-pass bf16 tensor cores (fast, ~1e-2 error:
 * opt-in only, does NOT meet the parity tolerance). */
enum { MDTB200_PREC_FP32 = 0, MDTB200_PREC_BF16X3 = 1, MDTB200_PREC_BF16 = 2 };

/* fused samplers (gc_sampling.py) */
enum { MDTB200_SAMPLER_DDIM = 0, MDTB200_SAMPLER_EULER = 1, MDTB200_SAMPLER_HEUN = 2, MDTB200_SAMPLER_DPMPP_2M = 3,
       MDTB200_SAMPLER_EULER_ANCESTRAL = 4 /* only through mdtb200_sample_ancestral (needs the caller's noise) */ };

/* goal modality: selects lang_emb vs goal_emb (mdtv_transformer.py:268-273) */
enum { MDTB200_MODALITY_VIS = 0, MDTB200_MODALITY_LANG = 1 };

typedef struct MdtConfig {
  int32_t abi_version;     /* = MDTB200_ABI_VERSION */
  int32_t variant;         /* MDTB200_VARIANT_* */
  int32_t embed_dim;       /* d: 384 (MDT-V) / 512 (MDT); multiple of 64 */
  int32_t n_heads;         /* 8; head_dim = d / n_heads must be <= 64 and a multiple of 4 */
  int32_t n_enc_layers;
  int32_t n_dec_layers;
  int32_t action_dim;      /* 7 */
  int32_t action_seq_len;  /* T_a = 10 (<= 16) */
  int32_t goal_dim;        /* 512 */
  int32_t obs_dim;         /* 384 (MDT-V) / 512 (MDT) */
  int32_t n_state_tokens;  /* MDT-V: obs_seq_len * n_obs_token = 3; MDT: 2 (static, gripper) */
  int32_t precision;       /* MDTB200_PREC_* */
  int32_t max_batch;       /* workspace is sized for this many samples per call */
  float   sigma_data;      /* 0.5 */
} MdtConfig;

typedef struct MdtHandle MdtHandle;

/* lifecycle ------------------------------------------------------------------------------ */
MDTB200_API int         mdtb200_abi_version(void);
MDTB200_API int         mdtb200_create(const MdtConfig* cfg, MdtHandle** out);
MDTB200_API void        mdtb200_destroy(MdtHandle* h);
/* message of the last failing call on this handle (h == NULL: last mdtb200_create failure) */
MDTB200_API const char* mdtb200_last_error(const MdtHandle* h);

/* weights: bind every tensor of the reference state dict by its key, e.g.
 * "inner_model.decoder.blocks.0.attn.key.weight" (same names/order the reference's EMA zip
 * relies on, mdt/models/mdtv_agent.py:152-158), then commit.  Unknown names are ignored
 * (pos_emb / proprio_emb are unused at inference in MDT-V); missing ones fail the commit. */
MDTB200_API int mdtb200_bind_weight(MdtHandle* h, const char* name, const float* dev_ptr, int64_t numel);
MDTB200_API int mdtb200_commit_weights(MdtHandle* h, void* stream);

/* forward_enc_only (mdtv_transformer.py:213-222): goal (B, goal_dim), state (B, n_state_tokens,
 * obs_dim) [MDT: token 0 = static, token 1 = gripper].  Stores the context and the decoder's
 * cross-attention K/V inside the handle; ctx_out (B, T_c, d) may be NULL.
 * goal_path_lang: MDTB200_MODALITY_*.  */
MDTB200_API int mdtb200_encode(MdtHandle* h, const float* goal, const float* state, int modality, int B,
                   float* ctx_out, void* stream);

/* forward_dec_only with an externally supplied context (mdtv_transformer.py:224-236):
 * loads ctx (B, T_c, d) into the handle and recomputes the cross-attention K/V. */
MDTB200_API int mdtb200_set_context(MdtHandle* h, const float* ctx, int B, void* stream);

/* One score-network evaluation on the cached context.
 *   precondition != 0: GCDenoiser.forward (score_wrappers.py:65-80):
 *        out = inner(x * c_in(sigma), sigma) * c_out(sigma) + x * c_skip(sigma)
 *   precondition == 0: raw MDTVTransformer.forward_dec_only(ctx, x, sigma).
 * x, out: (B, T_a, action_dim); sigma: (B,) per-sample noise levels (device). */
MDTB200_API int mdtb200_denoise(MdtHandle* h, const float* x, const float* sigma, int B, int precondition,
                    float* out, void* stream);

/* MDTVAgent.sample_loop for the fused samplers: encode once, then n_steps sampler iterations
 * (each 1 score evaluation; Heun 2 except on the last step), all inside one CUDA graph that
 * is cached per (B, n_steps, sampler, modality).
 *   sigmas: (n_steps + 1,) DEVICE array, last entry 0 (get_sigmas_*: gc_sampling.py:26-88)
 *   x_inout: (B, T_a, action_dim) device; holds x_T = randn * sigma_max on entry (drawn by the
 *            caller, mdtv_agent.py:546) and the sampled actions x_0 on return. */
MDTB200_API int mdtb200_sample(MdtHandle* h, int sampler, const float* sigmas, int n_steps,
                   const float* goal, const float* state, int modality, int B,
                   float* x_inout, void* stream);

/* Same call on HOST buffers (what a non-torch host binds): copies goal/state/x_T/sigmas to the
 * device, runs the graph, copies x_0 back and synchronises the stream before returning. */
MDTB200_API int mdtb200_sample_host(MdtHandle* h, int sampler, const float* sigmas_host, int n_steps,
                        const float* goal_host, const float* state_host, int modality, int B,
                        float* x_inout_host, void* stream);

/* sample_euler_ancestral (gc_sampling.py:213-253) as one CUDA graph.  The sampler is stochastic: `noise` holds the standard-normal
 * draws of every step, (n_steps, B, T_a, action_dim) on the device, produced by the CALLER in the reference's order (one
 * torch.randn_like per step whose sigma_down > 0, zeros elsewhere) so that the RNG stream is the reference's; eta as in the reference. */
MDTB200_API int mdtb200_sample_ancestral(MdtHandle* h, const float* sigmas, int n_steps, const float* goal, const float* state, int modality,
                                         int B, float* x_inout, const float* noise, float eta, void* stream);

/* introspection ---------------------------------------------------------------------------- */
/* number of kernels this handle has launched (graph replays count their kernel nodes) */
MDTB200_API int64_t mdtb200_launch_count(const MdtHandle* h);
/* copy a named internal buffer (debug / parity tests): "ctx", "kv", "mod", "xh", "a", "qkv",
 * "y", "h", "q" ... into dst (device, capacity in floats); returns #floats or <0 */
MDTB200_API int64_t mdtb200_debug_copy(MdtHandle* h, const char* name, float* dst_dev, int64_t capacity, void* stream);

/* training primitives --------------------------------------------------------------------------
 * Stateless exact-fp32 ops (no handle; errors through mdtb200_last_error(NULL)) from which the Python side builds the
 * autograd graph of GCDenoiser.loss (score_wrappers.py:45-63): every FLOP of the training forward and backward runs in
 * these kernels.  All tensors contiguous fp32 on the current device; `stream` as above. */
/* mode 0: C[M,N] = A[M,K] B[N,K]^T (+bias)   mode 1: C[M,K] = A[M,N] B[N,K]   mode 2: C[N,K] (+)= A[M,N]^T B[M,K] */
MDTB200_API int mdtb200_op_gemm(int mode, const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                int accumulate, void* stream);
/* the same products on tcgen05 (split-bf16 "bf16x3", fp32 accumulate): `scratch` = bf16 workspace of
 * mdtb200_op_gemm_tc_scratch(...) ELEMENTS; needs reduce / output-column dims that are multiples of 64 (else EUNSUPPORTED) */
MDTB200_API int64_t mdtb200_op_gemm_tc_scratch(int mode, int M, int N, int K);
MDTB200_API int mdtb200_op_gemm_tc(int mode, const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                   void* scratch, void* stream);
MDTB200_API int mdtb200_op_group_sum(const float* src, float* out, int G, int T, int C, int accumulate, void* stream);
MDTB200_API int mdtb200_op_colsum(const float* src, float* out, float* scratch, int M, int C, int accumulate, void* stream);
/* dy == NULL: out = act(x); else out = dy * act'(x).  act: 1 GELU(erf), 2 Mish, 3 SiLU */
MDTB200_API int mdtb200_op_act(const float* x, const float* dy, float* out, int64_t n, int act, void* stream);
MDTB200_API int mdtb200_op_ln_fwd(const float* x, const float* w, const float* b, const float* shift, const float* scale,
                                  int mod_stride, int rows_per_group, int M, int d, float* y, void* stream);
MDTB200_API int mdtb200_op_ln_bwd(const float* x, const float* dy, const float* w, const float* b, const float* scale,
                                  int mod_stride, int rows_per_group, int M, int d, float* dx, float* t_dw, float* t_db,
                                  float* t_dsc, void* stream);
/* p_drop / seed: dropout on the attention probabilities (transformer_blocks.py:142); the mask is a pure function of
 * (seed, element index), so the backward call with the same pair regenerates it */
MDTB200_API int mdtb200_op_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* y, int ldy,
                                    int B, int H, int hd, int Tq, int Tk, int causal, float p_drop, uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* dy, int lddy,
                                    float* dq, int lddq, float* dk, float* dv, int lddkv, int B, int H, int hd, int Tq, int Tk,
                                    int causal, float p_drop, uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_dropout(const float* x, float* out, int64_t n, float p, uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_gate_res(const float* x, const float* f, const float* gate, float* out, int M, int d,
                                    int rows_per_group, void* stream);
MDTB200_API int mdtb200_op_gate_res_bwd(const float* dout, const float* f, const float* gate, float* df, float* prod, int M, int d,
                                        int rows_per_group, void* stream);

/* PerceiverResampler ---------------------------------------------------------------------------------------
 * Replaces mdt/models/networks/transformers/perceiver_resampler.py:86-163 (PerceiverResampler.forward), the module that turns the
 * (B, T, n, d) Voltron token sequence into the num_latents state tokens the denoiser is conditioned on (mdtv_agent.py:392-403).
 * Weights are bound by the reference state-dict keys ("latents", "time_pos_emb", "layers.{l}.0.norm_media.weight", ...,
 * "layers.{l}.1.3.weight", "norm.weight"), then committed; forward runs on the stream passed in. */
typedef struct MdtPerceiverConfig {
  int32_t abi_version;      /* = MDTB200_ABI_VERSION */
  int32_t dim;              /* 384 */
  int32_t depth;            /* 6 (conf/model/mdtv_agent.yaml:28) */
  int32_t heads;            /* 8 */
  int32_t dim_head;         /* 64 */
  int32_t num_latents;      /* 3 = num_tokens_voltron; heads * num_latents <= 64 */
  int32_t num_time_embeds;  /* 1 */
  int32_t ff_mult;          /* 4 */
  int32_t max_batch;
  int32_t max_features;     /* T * n feature tokens per sample (392) */
} MdtPerceiverConfig;
typedef struct MdtPerceiver MdtPerceiver;
MDTB200_API int         mdtb200_perceiver_create(const MdtPerceiverConfig* cfg, MdtPerceiver** out);
MDTB200_API void        mdtb200_perceiver_destroy(MdtPerceiver* h);
MDTB200_API const char* mdtb200_perceiver_last_error(const MdtPerceiver* h);
MDTB200_API int         mdtb200_perceiver_bind_weight(MdtPerceiver* h, const char* name, const float* dev_ptr, int64_t numel);
MDTB200_API int         mdtb200_perceiver_commit_weights(MdtPerceiver* h, void* stream);
/* x_f (B, T, n, dim) device fp32; mask (B, T) device fp32 (1 = frame present) or NULL; out (B, num_latents, dim) */
MDTB200_API int         mdtb200_perceiver_forward(MdtPerceiver* h, const float* x_f, const float* mask, int B, int T, int n,
                                                  float* out, void* stream);
MDTB200_API int64_t     mdtb200_perceiver_launch_count(const MdtPerceiver* h);

/* Fused multi-tensor AdamW + EMA: one launch updates every parameter tensor of an optimizer group (torch.optim.AdamW's update
 * followed by the EMA callback's ema -= (1 - decay) (ema - w), mdt/callbacks/ema.py:117-126).  table: device array of
 * {float* param; const float* grad; float* exp_avg; float* exp_avg_sq; float* ema; int64_t numel; float step_size; float bc2_sqrt}
 * (step_size = lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t), t = that parameter's step count); blocks: device array of
 * int32 pairs {tensor index, 4096-element chunk index}, one per CUDA block.
 * step_dev (optional): device int32 step count; the bias corrections are then computed in the kernel (CUDA-graph capture). */
MDTB200_API int mdtb200_op_adamw_ema(const void* table, const void* blocks, int n_blocks, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, float ema_decay, int has_ema, const int* step_dev, void* stream);
/* device counter mixed into every dropout seed (NULL: off): fresh masks on every replay of a graph-captured training step */
MDTB200_API int mdtb200_op_set_seed_epoch(const uint64_t* dev_counter);

/* ---- training primitives, fused (round 2): every tensor is split ONCE into a row-major hi|lo bf16 operand ([rows, 2*cols]) that
 * serves the forward, input-gradient and weight-gradient GEMMs (the tcgen05 kernel reads it K-major or MN-major), the producers
 * (LayerNorm, attention, residual backward, activation backward) emit those operands directly, and the small reductions end in
 * per-CTA partial sums that one mdtb200_op_group_sum finishes.  Same reference ops as above (transformer_blocks.py). */
MDTB200_API int mdtb200_op_split(const float* x, const float* h, int act, void* out16, float* partial, int M, int K, void* stream);
MDTB200_API int mdtb200_op_split_rows_per_slab(void);
MDTB200_API int mdtb200_op_split_multi(const void* table, const void* blocks, int n_blocks, void* stream);
/* mode 0: C[M,N] = x16[M,2K] . w16[N,2K]^T + bias (epi 6: C = pre-activation, C16[M,2N] = split(GELU(C)));
 * mode 1: C[M,K] = dy16[M,2N] . w16[N,2K];  mode 2: C[N,K] = dy16[M,2N]^T . x16[M,2K], deterministic split-K over M when
 * splits > 1 (workspace of mdtb200_op_gemm16_ws floats, zero-initialised self-resetting counters, one per output tile) */
MDTB200_API int64_t mdtb200_op_gemm16_ws(int N, int K, int splits);
MDTB200_API int mdtb200_op_gemm16(int mode, const void* A16, const void* B16, const float* bias, float* C, void* C16, int M, int N, int K,
                                  int epi, int splits, float* sk_ws, unsigned* sk_cnt, void* stream);
MDTB200_API int mdtb200_op_ln_fwd16(const float* x, const float* w, const float* b, const float* shift, const float* scale, int mod_stride,
                                    int rows_per_group, int M, int d, float* y, void* y16, void* stream);
MDTB200_API int mdtb200_op_ln_bwd2_partials(int M, int T);   /* rows of the [.., 2d] partial buffer mdtb200_op_ln_bwd2 fills */
MDTB200_API int mdtb200_op_ln_bwd2(const float* x, const float* dy, const float* w, const float* b, const float* scale, int mod_stride,
                                   const float* dres, float* dx, float* dshift, float* dscale, int dmod_stride, float* partial, int M, int d,
                                   int T, void* stream);
MDTB200_API int mdtb200_op_attn_fwd16(const float* q, int ldq, const float* k, const float* v, int ldkv, void* y16, int B, int H, int hd,
                                      int Tq, int Tk, int causal, float p_drop, uint64_t seed, void* stream);
/* attention backward writing the projections' output gradients as GEMM operands (+ per-sample bias partials); shipped shapes only */
MDTB200_API int mdtb200_op_attn_bwd16(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* dy, int lddy,
                                      float* dk32, float* dv32, int ldkv32, void* dq16, int wq, int col0q, void* dkv16, int wkv, int col0k,
                                      int col0v, float* bpart, int bpart_kv, int B, int H, int hd, int Tq, int Tk, int causal, float p_drop,
                                      uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_res_drop_fwd(const float* x, const float* f, const float* gate, int gate_stride, float* out, int M, int d, int T,
                                        float p, uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_res_drop_bwd(const float* dout, const float* f, const float* gate, int gate_stride, void* df16, float* dgate,
                                        int dgate_stride, float* bpartial, int M, int d, int T, float p, uint64_t seed, void* stream);
MDTB200_API int mdtb200_op_narrow_fwd(const float* x, const float* W, const float* bias, float* y, int M, int K, int J, void* stream);
MDTB200_API int mdtb200_op_narrow_wgrad(const float* wide, const float* thin, float* partial, int M, int N, int J, int wide_major, void* stream);

/* Kernel timeline of everything the library launches (debugging / profiling aid, tools/ktrace.py): capacity > 0 arms the
 * trace, capacity == 0 copies up to max_records {globaltimer ns, tag|event|sm|grid|block} pairs to dst_host and disarms. */
MDTB200_API int64_t mdtb200_debug_ktrace(MdtHandle* h, int64_t capacity, unsigned long long* dst_host, int64_t max_records);

/* tests only: out (M,N) = epi(A (M,K) . W (N,K)^T + bias) through the tensor-core GEMM kernel on fp32 inputs
 * (epi: 0 none, 1 GELU, 4 residual R (M,N), 5 residual + gate[(m / rows_per_group), n] (gate row stride N)). */
MDTB200_API int mdtb200_debug_gemm(MdtHandle* h, const float* A, const float* W, const float* bias, const float* R,
                                   const float* gate, int M, int N, int K, int epi, int rows_per_group, float* out,
                                   void* stream);

/* tests / bench only: average duration (us) of `iters` back-to-back launches of the tensor-core GEMM kernel on zero
 * operands of the given shape (L2-warm), measured with CUDA events on `stream`. */
MDTB200_API int mdtb200_debug_gemm_time(MdtHandle* h, int M, int N, int K, int epi, int iters, float* avg_us, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDTB200_H_ */
