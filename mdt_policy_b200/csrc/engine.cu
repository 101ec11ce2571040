// libmdtb200.so -- host orchestration and C ABI (include/mdtb200.h) of the MDT denoising hot path.
//
// Data layout in HBM (all fp32 unless noted; M = B*T_a token rows, Mc = B*T_c context rows):
//   weights arena   packed copy of the reference state dict: per layer [Wq;Wk;Wv] (3d,d) fused,
//                   all decoder cross-attention [Wk;Wv] stacked (L*2d, d), all AdaLN modulation
//                   matrices stacked (L*6d, d); bf16 hi|lo split copies for the tensor-core path.
//   xh  (M, d)      decoder residual stream          xe (Mc, d)  encoder residual stream
//   a   (M, d)      LN(+modulate) output / GEMM A    qkv (M, 3d) y (M, d)   h (M, 4d)   q (M, d)
//   ctx (Mc, d)     encoder output                   kv (Mc, L*2d) cross-attention K|V of every decoder layer
//   mod (R, L*6d)   AdaLN shift/scale/gate table: R = n_steps rows when sampling (one sigma per step,
//                   shared by the batch) or R = B rows for per-sample sigma (denoise / loss API)
// One denoise step = a fixed kernel sequence (see decoder_eval); a full sampling call (encode + N steps)
// is captured once into a CUDA graph per (B, N, sampler, modality) and replayed.
#include "../../include/mdtb200.h"
#include "kernels_simt.cuh"
#include "gemm_tcgen05.cuh"
#include "fused_decoder.cuh"

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace mdt;

namespace {

thread_local std::string g_create_error;

struct EncLayerW {
  const float *ln1_w, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln2_w, *ln2_b, *wfc, *bfc, *wproj, *bproj;
  const __nv_bfloat16 *wqkv16, *wo16, *wfc16, *wproj16;
};
struct DecLayerW {
  const float *ln1_w, *ln1_b, *wqkv, *bqkv, *wo, *bo, *ln3_w, *ln3_b, *wq, *bq, *wco, *bco, *ln2_w, *ln2_b, *wfc, *bfc, *wproj, *bproj;
  const __nv_bfloat16 *wqkv16, *wo16, *wq16, *wco16, *wfc16, *wproj16;
  const __nv_bfloat16 *wgx16, *wux16;     // stacked per-head operands of the algebraic cross-attention tables (H*d, 2*64)
};
struct Weights {
  const float *goal0_w, *goal0_b, *goal2_w, *goal2_b;
  const float *lang0_w, *lang0_b, *lang2_w, *lang2_b;
  const __nv_bfloat16 *goal0_w16, *goal2_w16, *lang0_w16, *lang2_w16, *tok_w16, *incam_w16;
  const float *tok_w, *tok_b, *incam_w, *incam_b, *pos_emb;
  std::vector<EncLayerW> enc;
  std::vector<DecLayerW> dec;
  const float *enc_ln_w, *enc_ln_b, *dec_ln_w, *dec_ln_b;
  const float *wkv_all, *bkv_all, *wmod_all, *bmod_all, *bq_all;
  const __nv_bfloat16 *wkv_all16;
  const float *sig1_w, *sig1_b, *sig3_w, *sig3_b;
  const float *ae_w, *ae_b, *ap_w, *ap_b;
};

struct Bound { const float* ptr; int64_t numel; };

struct GraphKey {
  int B, n_steps, sampler, modality;
  bool operator<(const GraphKey& o) const {
    if (B != o.B) return B < o.B;
    if (n_steps != o.n_steps) return n_steps < o.n_steps;
    if (sampler != o.sampler) return sampler < o.sampler;
    return modality < o.modality;
  }
};
// device-side program of the persistent fused decoder (fused_decoder.cuh) for one (B, N, sampler) sampling call
struct FusedPlan {
  fd::Phase* d_prog = nullptr; fd::GemmDesc* d_gemms = nullptr; CUtensorMap* d_maps = nullptr; int n_phases = 0; int grid = 0; int n_rowgroups = 0;
  fd::FusedParams params{};
};
struct GraphEntry { cudaGraphExec_t exec; cudaGraphExec_t exec2; int64_t kernels; FusedPlan* plan; uint64_t last_use; };   // exec2: optional second segment

}  // namespace

// workspace pointers of one (sub-)batch; a sampling call splits the batch into independent branches (samples never
// interact), each with its own Work view, captured as parallel chains of the CUDA graph.
struct Work {
  float *in_goal, *in_state, *x, *x2, *dbuf, *gh, *xe, *ctx, *kv, *xh, *a, *qkv, *y, *hbuf, *q;
  __nv_bfloat16 *a16, *y16, *h16;
  // algebraic cross-attention (cross_row_kernel): head-major key / value operands and the G / U / c tables of layer 0
  __nv_bfloat16 *ka16, *va16; float *gtab, *utab, *ctab;
  // split-K workspace of the mlp c_proj GEMM (K = 4d): fp32 partial tiles + self-resetting per-tile counters, per sub-batch chain
  float* sk_ws; unsigned* sk_cnt;
  float* noise;           // ancestral sampler: per-step noise rows of this (sub-)batch, step stride = MdtHandle::noise_stride
};

struct MdtHandle {
  MdtConfig cfg;
  int device = 0;
  int d = 0, H = 0, hd = 0, T = 0, Tc = 0, Ts = 0, A = 0, Le = 0, Ld = 0;
  std::string err;
  int last_code = 0;
  std::map<std::string, Bound> bound;
  bool committed = false;
  int ctx_B = 0;                  // batch of the cached context (0 = none)

  float* arena = nullptr; size_t arena_floats = 0; size_t arena_used = 0;
  __nv_bfloat16* arena16 = nullptr; size_t arena16_elems = 0; size_t arena16_used = 0;
  Weights w;

  // workspace
  float *in_goal = nullptr, *in_state = nullptr, *x = nullptr, *x2 = nullptr, *dbuf = nullptr, *sigmas = nullptr;
  float *gh = nullptr, *xe = nullptr, *ctx = nullptr, *kv = nullptr, *xh = nullptr, *a = nullptr, *qkv = nullptr, *y = nullptr, *hbuf = nullptr, *q = nullptr;
  float *pe = nullptr, *sh = nullptr, *cs = nullptr, *mod = nullptr;
  __nv_bfloat16 *a16 = nullptr, *y16 = nullptr, *h16 = nullptr;   // split-bf16 operand copies (hi | lo)
  int mod_rows = 0;
  // algebraic cross-attention tables (see cross_row_kernel); per-layer strides in elements
  bool cross_fused = false;
  float* noise = nullptr; size_t noise_stride = 0; float* anc_eta = nullptr;     // fused euler_ancestral: static noise buffer (lazily allocated)
  static constexpr int SK_MAX_SPLITS = 4;
  float* sk_ws = nullptr; unsigned* sk_cnt = nullptr;
  int cproj_splits = 0;           // MDTB200_CPROJ_SPLITS: 0 = automatic (split-K when the mlp c_proj GEMMs of all chains leave SMs idle), 1..4 fixed
  int cur_chains = 1;             // concurrent sub-batch chains of the call being issued (sample_body)
  int pdl_late = 1;               // MDTB200_PDL_LATE=0 restores release-at-entry in the SIMT kernels (kernels_simt.cuh, pdl_enter_mode): +0.8 % at B = 256
  __nv_bfloat16 *ka16 = nullptr, *va16 = nullptr; float *gtab = nullptr, *utab = nullptr, *ctab = nullptr;
  size_t cross_rows = 0, ka_layer_stride = 0, tab_layer_stride = 0, ctab_layer_stride = 0;

  // persistent fused decoder (fused_decoder.cuh): split-K partial sums (C, Mp, d) and per-row-group progress counters
  bool fused = false; int fd_C = 0, fd_SPG = 0, fd_groups_max = 0;
  float* fd_partial = nullptr; int* fd_progress = nullptr; size_t fd_progress_bytes = 0;
  unsigned long long* fd_trace = nullptr;
  FusedPlan* cur_plan = nullptr;  // plan of the graph being captured
  uint64_t use_clock = 0;
  FusedPlan* denoise_plan = nullptr; int denoise_plan_B = 0, denoise_plan_pre = -1;

  cudaStream_t cap_stream = nullptr;
  static constexpr int MAX_BRANCHES = 8;
  static constexpr size_t MAX_GRAPHS = 16;
  cudaStream_t branch_streams[MAX_BRANCHES] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_BRANCHES] = {};
  int branches = 4;               // MDTB200_BRANCHES overrides

  Work work() const { return Work{in_goal, in_state, x, x2, dbuf, gh, xe, ctx, kv, xh, a, qkv, y, hbuf, q, a16, y16, h16, ka16, va16, gtab, utab, ctab, sk_ws, sk_cnt, noise}; }
  // view of samples [b0, ...): every buffer is row-indexed by sample (encoder rows reuse the decoder row offsets, Tc <= T);
  // mc_off = padded context rows of the sub-batches before this one (each sub-batch owns a 128-row aligned block per head)
  Work slice(const Work& w, int b0, int mc_off = 0) const {
    const size_t r = (size_t)b0 * T, dd = d;
    Work o = w;
    o.in_goal += (size_t)b0 * cfg.goal_dim; o.in_state += (size_t)b0 * Ts * cfg.obs_dim;
    o.x += r * A; o.x2 += r * A; o.dbuf += r * A; o.gh += (size_t)b0 * 2 * dd;
    if (o.noise) o.noise += r * A;
    o.xe += r * dd; o.ctx += (size_t)b0 * Tc * dd; o.kv += (size_t)b0 * Tc * Ld * 2 * dd;
    o.xh += r * dd; o.a += r * dd; o.qkv += r * 3 * dd; o.y += r * dd; o.hbuf += r * 4 * dd; o.q += r * dd;
    if (o.a16) { o.a16 += r * 2 * dd; o.y16 += r * 2 * dd; o.h16 += r * 8 * dd; }
    if (o.sk_ws) {      // disjoint m-tile ranges per chain: floor(r / 128) + floor(b0 / 32) (chains have >= 32 samples)
      const size_t t0 = r / 128 + (size_t)b0 / 32;
      o.sk_ws += t0 * SK_MAX_SPLITS * 128 * dd; o.sk_cnt += t0 * (dd / 64);
    }
    if (o.gtab) {
      const size_t hr = (size_t)H * mc_off;
      o.ka16 += hr * 128; o.va16 += hr * 128; o.gtab += hr * dd; o.utab += hr * dd; o.ctab += (size_t)b0 * H * Tc;
    }
    return o;
  }
  std::map<GraphKey, GraphEntry> graphs;
  int64_t launches = 0;
  int64_t capture_count = 0;      // kernels launched while capturing
  bool capturing = false;
  std::vector<void*> allocs;
  tc::TmaEncoder tma;
};

namespace {

int fail(MdtHandle* h, int code, const char* fmt, ...) {
  if (h) h->last_code = code;
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}

#define CUDA_TRY(h, expr)                                                                             \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess) return fail(h, MDTB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline void count_launch(MdtHandle* h) { if (h->capturing) h->capture_count++; else h->launches++; }

inline int check_launch(MdtHandle* h, const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, MDTB200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ launchers

struct Gemm {
  const float* A = nullptr; int lda = 0;                 // fp32 A (SIMT path)
  const __nv_bfloat16* A16 = nullptr; int lda16 = 0;     // split-bf16 A (hi | lo at +K) for the tensor-core path
  const float* W = nullptr; const __nv_bfloat16* W16 = nullptr; const float* bias = nullptr;
  float* C = nullptr; int ldc = 0;
  __nv_bfloat16* C16 = nullptr; int ldc16 = 0; int lo_off = 0;
  const float* R = nullptr; int ldr = 0;
  const float* gate = nullptr; int gate_stride = 0; int rows_per_group = 1;
  int M = 0, N = 0, K = 0; int epi = EPI_NONE;
  int gi = 0, go = 0, goff = 0;
  int wg_rows = 0, wg_stride = 0, w_rows = 0;          // weight groups (tensor-core path only)
  int splits = 1; float* sk_ws = nullptr; unsigned* sk_cnt = nullptr;     // deterministic split-K (tensor-core path only)
};

int launch_sgemm(MdtHandle* h, const Gemm& p, cudaStream_t st) {
  if (p.M <= SKINNY_MAXM && p.gi == 0 && !p.C16 && p.epi <= EPI_SILU && p.lda == p.K && p.ldc == p.N && p.K % 4 == 0 &&
      (size_t)p.M * p.K * 4 <= 48 * 1024) {
    SkinnyArgs g{p.A, p.W, p.bias, p.C, p.M, p.N, p.K, p.epi};
    launch_pdl(skinny_gemm_kernel, dim3((p.N + 7) / 8), dim3(256), (size_t)p.M * p.K * 4, st, g);
    count_launch(h);
    return check_launch(h, "skinny_gemm_kernel");
  }
  if (p.K % SG_BK != 0 || p.N % 4 != 0 || p.lda % 4 != 0) return fail(h, MDTB200_EINVAL, "sgemm: unsupported shape M=%d N=%d K=%d", p.M, p.N, p.K);
  GemmArgs g{};
  g.A = p.A; g.lda = p.lda; g.W = p.W; g.bias = p.bias; g.C = p.C; g.ldc = p.ldc; g.R = p.R; g.ldr = p.ldr;
  g.gate = p.gate; g.gate_stride = p.gate_stride; g.rows_per_group = p.rows_per_group;
  g.M = p.M; g.N = p.N; g.K = p.K; g.gi = p.gi; g.go = p.go; g.goff = p.goff;
  g.C16 = p.C16; g.ldc16 = p.ldc16; g.lo_off = p.lo_off;
  dim3 grid((p.N + SG_BN - 1) / SG_BN, (p.M + SG_BM - 1) / SG_BM);
  switch (p.epi) {
    case EPI_NONE: launch_pdl(sgemm_tn_kernel<EPI_NONE>, grid, dim3(SG_THREADS), 0, st, g); break;
    case EPI_GELU: launch_pdl(sgemm_tn_kernel<EPI_GELU>, grid, dim3(SG_THREADS), 0, st, g); break;
    case EPI_MISH: launch_pdl(sgemm_tn_kernel<EPI_MISH>, grid, dim3(SG_THREADS), 0, st, g); break;
    case EPI_SILU: launch_pdl(sgemm_tn_kernel<EPI_SILU>, grid, dim3(SG_THREADS), 0, st, g); break;
    case EPI_RES: launch_pdl(sgemm_tn_kernel<EPI_RES>, grid, dim3(SG_THREADS), 0, st, g); break;
    case EPI_RES_GATE: launch_pdl(sgemm_tn_kernel<EPI_RES_GATE>, grid, dim3(SG_THREADS), 0, st, g); break;
    default: return fail(h, MDTB200_EINVAL, "sgemm: bad epilogue %d", p.epi);
  }
  count_launch(h);
  return check_launch(h, "sgemm_tn_kernel");
}

// GEMM dispatcher: tensor cores when the handle's precision asks for them and the operand is available in
// split-bf16 form, exact fp32 CUDA cores otherwise (tiny GEMMs of the sigma path always stay fp32).
int gemm(MdtHandle* h, const Gemm& p, cudaStream_t st) {
  if (h->cfg.precision != MDTB200_PREC_FP32 && p.A16 && p.W16) {
    tc::TcGemm t{};
    t.A16 = p.A16; t.lda16 = p.lda16; t.W16 = p.W16; t.bias = p.bias;
    t.C = p.C; t.ldc = p.ldc; t.C16 = p.C16; t.ldc16 = p.ldc16; t.lo_off = p.lo_off;
    t.R = p.R; t.ldr = p.ldr; t.gate = p.gate; t.gate_stride = p.gate_stride; t.rows_per_group = p.rows_per_group;
    t.M = p.M; t.N = p.N; t.K = p.K; t.epi = p.epi;
    t.passes = h->cfg.precision == MDTB200_PREC_BF16X3 ? 3 : 1;
    t.trace = nullptr; t.gi = p.gi; t.go = p.go; t.goff = p.goff;
    t.wg_rows = p.wg_rows; t.wg_stride = p.wg_stride; t.w_rows = p.w_rows;
    if (p.splits > 1 && p.sk_ws && t.passes == 3) { t.splits = p.splits; t.sk_ws = p.sk_ws; t.sk_cnt = p.sk_cnt; }
    const char* e = tc::launch_tc_gemm(h->tma, t, st);
    if (e) return fail(h, MDTB200_ECUDA, "tcgen05 gemm (M=%d N=%d K=%d): %s", p.M, p.N, p.K, e);
    count_launch(h);
    return check_launch(h, "tc_gemm_kernel");
  }
  if (!p.A) return fail(h, MDTB200_EINVAL, "gemm: no fp32 operand for the CUDA-core path");
  return launch_sgemm(h, p, st);
}

int launch_ln(MdtHandle* h, const float* x, float* out, __nv_bfloat16* out16, const float* w, const float* b,
              const float* shift, const float* scale, int mod_stride, int M, cudaStream_t st) {
  LnArgs a{};
  a.x = x; a.out = out; a.out16 = out16; a.ld16 = 2 * h->d; a.lo_off = h->d; a.w = w; a.b = b; a.shift = shift; a.scale = scale;
  a.mod_stride = mod_stride; a.rows_per_group = h->T; a.M = M; a.d = h->d; a.late = h->pdl_late;
  int blocks = (M * 32 + 255) / 256;
  switch (h->d / 128) {
    case 1: launch_pdl(ln_mod_kernel<1>, dim3(blocks), dim3(256), 0, st, a); break;
    case 2: launch_pdl(ln_mod_kernel<2>, dim3(blocks), dim3(256), 0, st, a); break;
    case 3: launch_pdl(ln_mod_kernel<3>, dim3(blocks), dim3(256), 0, st, a); break;
    case 4: launch_pdl(ln_mod_kernel<4>, dim3(blocks), dim3(256), 0, st, a); break;
    case 6: launch_pdl(ln_mod_kernel<6>, dim3(blocks), dim3(256), 0, st, a); break;
    case 8: launch_pdl(ln_mod_kernel<8>, dim3(blocks), dim3(256), 0, st, a); break;
    default: return fail(h, MDTB200_EUNSUPPORTED, "embed_dim %d not supported by ln kernel", h->d);
  }
  count_launch(h);
  return check_launch(h, "ln_mod_kernel");
}

int launch_attn(MdtHandle* h, const float* q, int ldq, const float* k, const float* v, int ldkv, float* y, __nv_bfloat16* y16,
                int B, int Tq, int Tk, int causal, cudaStream_t st) {
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.k = k; a.v = v; a.ldkv = ldkv; a.y = y; a.ldy = h->d; a.y16 = y16; a.ld16 = 2 * h->d; a.lo_off = h->d;
  a.B = B; a.H = h->H; a.hd = h->hd; a.Tq = Tq; a.Tk = Tk; a.causal = causal;
  a.scale = 1.0f / sqrtf((float)h->hd); a.late = h->pdl_late;
  // shipped shapes run the compile-time specialised kernel (2 heads per CTA); anything else the generic one
  const bool c = causal != 0;
  #define ATT_CASE(HD, TQ, TK, CA)                                                                                  \
    if (h->hd == HD && Tq == TQ && Tk == TK && c == (CA != 0) && h->H % 2 == 0) {                                   \
      launch_pdl(attention_fixed_kernel<HD, TQ, TK, CA, 2>, dim3(B, h->H / 2), dim3(128), 0, st, a);                \
      count_launch(h);                                                                                              \
      return check_launch(h, "attention_fixed_kernel");                                                             \
    }
  ATT_CASE(48, 10, 10, 1) ATT_CASE(48, 10, 4, 1) ATT_CASE(48, 4, 4, 0)
  ATT_CASE(64, 10, 10, 1) ATT_CASE(64, 10, 3, 1) ATT_CASE(64, 3, 3, 0)
  #undef ATT_CASE
  launch_pdl(attention_kernel, dim3(B), dim3(ATT_THREADS), attention_smem_bytes(h->d, h->H, Tq, Tk), st, a);
  count_launch(h);
  return check_launch(h, "attention_kernel");
}

int launch_head(MdtHandle* h, HeadArgs a, cudaStream_t st) {
  int blocks = (a.M * 32 + 255) / 256;
  switch (h->d / 128) {
    case 1: launch_pdl(head_kernel<1>, dim3(blocks), dim3(256), 0, st, a); break;
    case 2: launch_pdl(head_kernel<2>, dim3(blocks), dim3(256), 0, st, a); break;
    case 3: launch_pdl(head_kernel<3>, dim3(blocks), dim3(256), 0, st, a); break;
    case 4: launch_pdl(head_kernel<4>, dim3(blocks), dim3(256), 0, st, a); break;
    case 6: launch_pdl(head_kernel<6>, dim3(blocks), dim3(256), 0, st, a); break;
    case 8: launch_pdl(head_kernel<8>, dim3(blocks), dim3(256), 0, st, a); break;
    default: return fail(h, MDTB200_EUNSUPPORTED, "embed_dim %d not supported by head kernel", h->d);
  }
  count_launch(h);
  return check_launch(h, "head_kernel");
}

#define TRY(expr) do { int rc_ = (expr); if (rc_) return rc_; } while (0)

inline bool use_tc(const MdtHandle* h) { return h->cfg.precision != MDTB200_PREC_FP32; }

// ------------------------------------------------------------------------------------------ network pieces

// number of concurrent sub-batch chains for a sampling call: sub-batches of >= 32 samples, rows a multiple of 128 if possible
int branch_count(const MdtHandle* h, int B) {
  int nb = h->branches < 1 ? 1 : (h->branches > MdtHandle::MAX_BRANCHES ? MdtHandle::MAX_BRANCHES : h->branches);
  while (nb > 1 && B / nb < 32) --nb;
  return nb;
}

inline int round128(int v) { return (v + 127) / 128 * 128; }

// G / U / c tables of the algebraic cross-attention for the B samples of this (sub-)batch (kernels_simt.cuh, cross_row_kernel):
// one packing kernel for all layers, then two head-grouped tensor-core GEMMs per layer (K = head_dim padded to 64).
int compute_cross_tables(MdtHandle* h, const Work& k, int B, cudaStream_t st) {
  const int d = h->d, H = h->H, Tc = h->Tc, Mc = B * Tc, mcp = round128(Mc), Ld = h->Ld;
  CrossPackArgs a{};
  a.kv = k.kv; a.ldkv = Ld * 2 * d; a.bq_all = h->w.bq_all; a.ka = k.ka16; a.va = k.va16; a.layer_stride16 = h->ka_layer_stride;
  a.ctab = k.ctab; a.ctab_layer_stride = h->ctab_layer_stride; a.Mc = Mc; a.mcp = mcp; a.Tc = Tc; a.H = H; a.hd = h->hd; a.d = d; a.L = Ld;
  const int n = H * mcp * 64;
  launch_pdl(pack_cross_operands_kernel, dim3((n + 255) / 256, Ld), dim3(256), 0, st, a);
  count_launch(h);
  TRY(check_launch(h, "pack_cross_operands_kernel"));
  for (int l = 0; l < Ld; ++l) {
    for (int which = 0; which < 2; ++which) {
      Gemm g;
      g.A16 = (which ? k.va16 : k.ka16) + (size_t)l * h->ka_layer_stride; g.lda16 = 128;
      g.W16 = which ? h->w.dec[l].wux16 : h->w.dec[l].wgx16; g.w_rows = H * d; g.wg_rows = mcp; g.wg_stride = d;
      g.C = (which ? k.utab : k.gtab) + (size_t)l * h->tab_layer_stride; g.ldc = d;
      g.M = H * mcp; g.N = d; g.K = 64;
      TRY(gemm(h, g, st));
    }
  }
  return 0;
}

// cross-attention K/V of every decoder layer from the context: kv[Mc, L*2d] = ctx . Wkv_all^T + b
int compute_kv(MdtHandle* h, const Work& k, int B, cudaStream_t st, bool ctx_split_valid = false) {
  Gemm g;
  g.A = k.ctx; g.lda = h->d; g.W = h->w.wkv_all; g.bias = h->w.bkv_all; g.C = k.kv; g.ldc = h->Ld * 2 * h->d;
  g.M = B * h->Tc; g.N = h->Ld * 2 * h->d; g.K = h->d;
  if (ctx_split_valid) { g.A16 = k.a16; g.lda16 = 2 * h->d; g.W16 = h->w.wkv_all16; }
  TRY(gemm(h, g, st));
  if (h->cross_fused) TRY(compute_cross_tables(h, k, B, st));
  return 0;
}

// forward_enc_only (mdtv_transformer.py:213-222; mdt_transformer.py:211-229 for the MDT variant)
int encoder(MdtHandle* h, const Work& k, const float* goal, const float* state, int modality, int B, cudaStream_t st) {
  const int d = h->d, Tc = h->Tc, Ts = h->Ts, Mc = B * Tc;
  const Weights& w = h->w;
  const bool lang = modality == MDTB200_MODALITY_LANG && w.lang0_w != nullptr;
  const bool tcp = use_tc(h);
  const int G = h->cfg.goal_dim, O = h->cfg.obs_dim;
  // split-bf16 staging of the embedding inputs inside the (idle) h16 buffer: goal | state | goal-MLP hidden
  __nv_bfloat16* g16 = k.h16;
  __nv_bfloat16* s16 = tcp ? k.h16 + (size_t)B * 2 * G : nullptr;
  __nv_bfloat16* gh16 = tcp ? s16 + (size_t)B * Ts * 2 * O : nullptr;
  if (tcp) {
    if ((size_t)B * (2 * G + Ts * 2 * O + 4 * d) > (size_t)B * h->T * 8 * d) return fail(h, MDTB200_EUNSUPPORTED, "embedding staging does not fit");
    int64_t n = (int64_t)B * G;
    launch_pdl(split_weights_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, goal, g16, (int64_t)B, G);
    count_launch(h);
    n = (int64_t)B * Ts * O;
    launch_pdl(split_weights_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, state, s16, (int64_t)B * Ts, O);
    count_launch(h);
    TRY(check_launch(h, "split_weights_kernel"));
  }
  {  // goal MLP: Linear(goal_dim, 2d) -> GELU -> Linear(2d, d), written to context token 0
    Gemm g;
    g.A = goal; g.lda = G; g.W = lang ? w.lang0_w : w.goal0_w; g.bias = lang ? w.lang0_b : w.goal0_b;
    g.C = tcp ? nullptr : k.gh; g.ldc = 2 * d; g.M = B; g.N = 2 * d; g.K = G; g.epi = EPI_GELU;
    if (tcp) { g.A16 = g16; g.lda16 = 2 * G; g.W16 = lang ? w.lang0_w16 : w.goal0_w16; g.C16 = gh16; g.ldc16 = 4 * d; g.lo_off = 2 * d; }
    TRY(gemm(h, g, st));
    Gemm g2;
    g2.A = k.gh; g2.lda = 2 * d; g2.W = lang ? w.lang2_w : w.goal2_w; g2.bias = lang ? w.lang2_b : w.goal2_b;
    g2.C = k.xe; g2.ldc = d; g2.M = B; g2.N = d; g2.K = 2 * d; g2.gi = 1; g2.go = Tc; g2.goff = 0;
    if (tcp) { g2.A16 = gh16; g2.lda16 = 4 * d; g2.W16 = lang ? w.lang2_w16 : w.goal2_w16; }
    TRY(gemm(h, g2, st));
  }
  if (h->cfg.variant == MDTB200_VARIANT_MDTV) {  // tok_emb on the n_state_tokens Voltron tokens -> context tokens 1..
    Gemm g;
    g.A = state; g.lda = O; g.W = w.tok_w; g.bias = w.tok_b; g.C = k.xe; g.ldc = d;
    g.M = B * Ts; g.N = d; g.K = O; g.gi = Ts; g.go = Tc; g.goff = 1;
    if (tcp) { g.A16 = s16; g.lda16 = 2 * O; g.W16 = w.tok_w16; }
    TRY(gemm(h, g, st));
  } else {  // MDT: token 1 = tok_emb(static), token 2 = incam_embed(gripper); then learned pos_emb
    for (int t = 0; t < 2; ++t) {
      Gemm g;
      g.A = state + (size_t)t * O; g.lda = 2 * O; g.W = t == 0 ? w.tok_w : w.incam_w; g.bias = t == 0 ? w.tok_b : w.incam_b;
      g.C = k.xe; g.ldc = d; g.M = B; g.N = d; g.K = O; g.gi = 1; g.go = Tc; g.goff = 1 + t;
      if (tcp) { g.A16 = s16 + (size_t)t * 2 * O; g.lda16 = 4 * O; g.W16 = t == 0 ? w.tok_w16 : w.incam_w16; }
      TRY(gemm(h, g, st));
    }
    int n = Mc * d;
    launch_pdl(add_pos_emb_kernel, dim3((n + 255) / 256), dim3(256), 0, st, k.xe, w.pos_emb, B, Tc, d);
    count_launch(h);
    TRY(check_launch(h, "add_pos_emb_kernel"));
  }
  for (int l = 0; l < h->Le; ++l) {  // Block.forward, transformer_blocks.py:209-214
    const EncLayerW& L = w.enc[l];
    TRY(launch_ln(h, k.xe, tcp ? nullptr : k.a, tcp ? k.a16 : nullptr, L.ln1_w, L.ln1_b, nullptr, nullptr, 0, Mc, st));
    Gemm g;
    g.A = k.a; g.lda = d; g.A16 = k.a16; g.lda16 = 2 * d; g.W = L.wqkv; g.W16 = L.wqkv16; g.bias = L.bqkv; g.C = k.qkv; g.ldc = 3 * d;
    g.M = Mc; g.N = 3 * d; g.K = d;
    TRY(gemm(h, g, st));
    TRY(launch_attn(h, k.qkv, 3 * d, k.qkv + d, k.qkv + 2 * d, 3 * d, tcp ? nullptr : k.y, tcp ? k.y16 : nullptr, B, Tc, Tc, 0, st));
    Gemm o;
    o.A = k.y; o.lda = d; o.A16 = k.y16; o.lda16 = 2 * d; o.W = L.wo; o.W16 = L.wo16; o.bias = L.bo; o.C = k.xe; o.ldc = d; o.R = k.xe; o.ldr = d;
    o.M = Mc; o.N = d; o.K = d; o.epi = EPI_RES;
    TRY(gemm(h, o, st));
    TRY(launch_ln(h, k.xe, tcp ? nullptr : k.a, tcp ? k.a16 : nullptr, L.ln2_w, L.ln2_b, nullptr, nullptr, 0, Mc, st));
    Gemm f;
    f.A = k.a; f.lda = d; f.A16 = k.a16; f.lda16 = 2 * d; f.W = L.wfc; f.W16 = L.wfc16; f.bias = L.bfc;
    f.C = tcp ? nullptr : k.hbuf; f.ldc = 4 * d; f.C16 = tcp ? k.h16 : nullptr; f.ldc16 = 8 * d; f.lo_off = 4 * d;
    f.M = Mc; f.N = 4 * d; f.K = d; f.epi = EPI_GELU;
    TRY(gemm(h, f, st));
    Gemm p;
    p.A = k.hbuf; p.lda = 4 * d; p.A16 = k.h16; p.lda16 = 8 * d; p.W = L.wproj; p.W16 = L.wproj16; p.bias = L.bproj; p.C = k.xe; p.ldc = d; p.R = k.xe; p.ldr = d;
    p.M = Mc; p.N = d; p.K = 4 * d; p.epi = EPI_RES;
    TRY(gemm(h, p, st));
  }
  TRY(launch_ln(h, k.xe, k.ctx, tcp ? k.a16 : nullptr, w.enc_ln_w, w.enc_ln_b, nullptr, nullptr, 0, Mc, st));
  TRY(compute_kv(h, k, B, st, tcp));
  h->ctx_B = B;
  return 0;
}

// AdaLN table: mod[r, :] for R sigma values (process_sigma_embeddings :238-244, sigma_emb :169-174,
// AdaLNZero :245-260 of every decoder layer, stacked).
int sigma_path(MdtHandle* h, const float* sigma, int R, cudaStream_t st) {
  const int d = h->d;
  int n = R * (d / 2);
  launch_pdl(sigma_posemb_kernel, dim3((n + 127) / 128), dim3(128), 0, st, sigma, R, d, h->pe);
  count_launch(h);
  TRY(check_launch(h, "sigma_posemb_kernel"));
  Gemm g1;
  g1.A = h->pe; g1.lda = d; g1.W = h->w.sig1_w; g1.bias = h->w.sig1_b; g1.C = h->sh; g1.ldc = 2 * d; g1.M = R; g1.N = 2 * d; g1.K = d; g1.epi = EPI_MISH;
  TRY(gemm(h, g1, st));
  Gemm g2;   // epilogue applies the SiLU of AdaLNZero.modulation[0] directly: c itself is never needed
  g2.A = h->sh; g2.lda = 2 * d; g2.W = h->w.sig3_w; g2.bias = h->w.sig3_b; g2.C = h->cs; g2.ldc = d; g2.M = R; g2.N = d; g2.K = 2 * d; g2.epi = EPI_SILU;
  TRY(gemm(h, g2, st));
  Gemm g3;
  g3.A = h->cs; g3.lda = d; g3.W = h->w.wmod_all; g3.bias = h->w.bmod_all; g3.C = h->mod; g3.ldc = h->Ld * 6 * d; g3.M = R; g3.N = h->Ld * 6 * d; g3.K = d;
  TRY(gemm(h, g3, st));
  return 0;
}

// One score-network evaluation (forward_dec_only :224-236 + ConditionedBlock :292-309) on the cached context.
//   x_in     actions the network sees (before c_in scaling)
//   mod      AdaLN rows for this evaluation; mod_stride = 0 -> one row shared by the batch
//   sigma/sigma_stride   per-sample sigma for the c_in scaling (stride 0 -> shared)
int decoder_eval(MdtHandle* h, const Work& k, const float* x_in, const float* mod, int mod_stride, const float* sigma, int sigma_stride,
                 int precondition, int B, HeadArgs head, cudaStream_t st) {
  const int d = h->d, T = h->T, Tc = h->Tc, M = B * T;
  const Weights& w = h->w;
  const bool tcp = use_tc(h);
  {
    ActEmbArgs a{};
    a.x = x_in; a.sigma = sigma; a.sigma_stride = sigma_stride; a.T = T; a.A = h->A; a.d = d; a.M = M;
    a.W = w.ae_w; a.b = w.ae_b; a.xh = k.xh; a.sigma_data = h->cfg.sigma_data; a.precondition = precondition;
    int n = M * d;
    launch_pdl(action_embed_kernel, dim3((n + 255) / 256), dim3(256), 0, st, a);
    count_launch(h);
    TRY(check_launch(h, "action_embed_kernel"));
  }
  const int kvld = h->Ld * 2 * d;
  for (int l = 0; l < h->Ld; ++l) {
    const DecLayerW& L = w.dec[l];
    const float* ml = mod + (size_t)l * 6 * d;   // shift_msa | scale_msa | gate_msa | shift_mlp | scale_mlp | gate_mlp
    // x += gate_msa * SelfAttn_causal(shift_msa + LN1(x) * scale_msa)
    TRY(launch_ln(h, k.xh, tcp ? nullptr : k.a, tcp ? k.a16 : nullptr, L.ln1_w, L.ln1_b, ml, ml + d, mod_stride, M, st));
    Gemm g;
    g.A = k.a; g.lda = d; g.A16 = k.a16; g.lda16 = 2 * d; g.W = L.wqkv; g.W16 = L.wqkv16; g.bias = L.bqkv; g.C = k.qkv; g.ldc = 3 * d; g.M = M; g.N = 3 * d; g.K = d;
    TRY(gemm(h, g, st));
    TRY(launch_attn(h, k.qkv, 3 * d, k.qkv + d, k.qkv + 2 * d, 3 * d, tcp ? nullptr : k.y, tcp ? k.y16 : nullptr, B, T, T, 1, st));
    Gemm o;
    o.A = k.y; o.lda = d; o.A16 = k.y16; o.lda16 = 2 * d; o.W = L.wo; o.W16 = L.wo16; o.bias = L.bo; o.C = k.xh; o.ldc = d; o.R = k.xh; o.ldr = d;
    o.gate = ml + 2 * d; o.gate_stride = mod_stride; o.rows_per_group = T; o.M = M; o.N = d; o.K = d; o.epi = EPI_RES_GATE;
    TRY(gemm(h, o, st));
    // x += CrossAttn_causal-top-left(LN3(x), ctx)       (ln3 is nn.LayerNorm with bias)
    if (h->cross_fused) {
      // algebraic form on the per-call G / U tables: LN3 -> scores -> softmax -> output projection -> residual -> LN2 + modulate
      CrossRowArgs ca{};
      ca.xh = k.xh; ca.gtab = k.gtab + (size_t)l * h->tab_layer_stride; ca.utab = k.utab + (size_t)l * h->tab_layer_stride;
      ca.ctab = k.ctab + (size_t)l * h->ctab_layer_stride; ca.mcp = round128(B * Tc);
      ca.ln3_w = L.ln3_w; ca.ln3_b = L.ln3_b; ca.bco = L.bco; ca.ln2_w = L.ln2_w; ca.ln2_b = L.ln2_b;
      ca.shift = ml + 3 * d; ca.scale = ml + 4 * d; ca.mod_stride = mod_stride;
      ca.a16 = k.a16; ca.ld16 = 2 * d; ca.lo_off = d; ca.B = B; ca.T = T; ca.Tc = Tc; ca.H = h->H; ca.d = d;
      {   // the predecessor is the tcgen05 O GEMM (cross_fused implies the tensor-core path): tables / parameters may be read early
        static const int early = getenv("MDTB200_CROSS_EARLY") ? atoi(getenv("MDTB200_CROSS_EARLY")) : 1;
        ca.early = tcp && early; ca.late = h->pdl_late;
      }
      const size_t smem = cross_row_smem_bytes(d, T, Tc, h->H);
      if (d == 384) launch_pdl(cross_row_kernel<3>, dim3(B), dim3(CR_THREADS), smem, st, ca);
      else launch_pdl(cross_row_kernel<4>, dim3(B), dim3(CR_THREADS), smem, st, ca);
      count_launch(h);
      TRY(check_launch(h, "cross_row_kernel"));
    } else {
    TRY(launch_ln(h, k.xh, tcp ? nullptr : k.a, tcp ? k.a16 : nullptr, L.ln3_w, L.ln3_b, nullptr, nullptr, 0, M, st));
    Gemm cq;
    cq.A = k.a; cq.lda = d; cq.A16 = k.a16; cq.lda16 = 2 * d; cq.W = L.wq; cq.W16 = L.wq16; cq.bias = L.bq; cq.C = k.q; cq.ldc = d; cq.M = M; cq.N = d; cq.K = d;
    TRY(gemm(h, cq, st));
    TRY(launch_attn(h, k.q, d, k.kv + (size_t)l * 2 * d, k.kv + (size_t)l * 2 * d + d, kvld, tcp ? nullptr : k.y, tcp ? k.y16 : nullptr, B, T, Tc, 1, st));
    Gemm co;
    co.A = k.y; co.lda = d; co.A16 = k.y16; co.lda16 = 2 * d; co.W = L.wco; co.W16 = L.wco16; co.bias = L.bco; co.C = k.xh; co.ldc = d; co.R = k.xh; co.ldr = d;
    co.M = M; co.N = d; co.K = d; co.epi = EPI_RES;
    TRY(gemm(h, co, st));
    // x += gate_mlp * MLP(shift_mlp + LN2(x) * scale_mlp)
    TRY(launch_ln(h, k.xh, tcp ? nullptr : k.a, tcp ? k.a16 : nullptr, L.ln2_w, L.ln2_b, ml + 3 * d, ml + 4 * d, mod_stride, M, st));
    }
    Gemm f;
    f.A = k.a; f.lda = d; f.A16 = k.a16; f.lda16 = 2 * d; f.W = L.wfc; f.W16 = L.wfc16; f.bias = L.bfc;
    f.C = tcp ? nullptr : k.hbuf; f.ldc = 4 * d; f.C16 = tcp ? k.h16 : nullptr; f.ldc16 = 8 * d; f.lo_off = 4 * d;
    f.M = M; f.N = 4 * d; f.K = d; f.epi = EPI_GELU;
    TRY(gemm(h, f, st));
    Gemm p;
    p.A = k.hbuf; p.lda = 4 * d; p.A16 = k.h16; p.lda16 = 8 * d; p.W = L.wproj; p.W16 = L.wproj16; p.bias = L.bproj; p.C = k.xh; p.ldc = d; p.R = k.xh; p.ldr = d;
    p.gate = ml + 5 * d; p.gate_stride = mod_stride; p.rows_per_group = T; p.M = M; p.N = d; p.K = 4 * d; p.epi = EPI_RES_GATE;
    // K = 4d on a d-wide output is the longest serial loop of the layer: split it when the chains together leave SMs idle
    // (small batches: -7 % per call at B = 1, -4.5 % at B = 64; at B = 256 the four chains fill the machine and splitting costs 5 %)
    int sp = h->cproj_splits;
    if (sp == 0) {
      const int ctas = h->cur_chains * ((M + 127) / 128) * (d / 64);
      sp = ctas * 4 <= 160 ? 4 : ctas * 3 <= 160 ? 3 : ctas * 2 <= 160 ? 2 : 1;
    }
    p.splits = sp; p.sk_ws = k.sk_ws; p.sk_cnt = k.sk_cnt;
    TRY(gemm(h, p, st));
  }
  head.xh = k.xh; head.lnw = w.dec_ln_w; head.lnb = w.dec_ln_b; head.W = w.ap_w; head.bias = w.ap_b;
  head.M = M; head.d = d; head.A = h->A; head.T = T; head.sigma_data = h->cfg.sigma_data;
  return launch_head(h, head, st);
}

// sampler iterations of one branch (sub-batch) on its own stream
int sample_steps(MdtHandle* h, const Work& k, int sampler, int n_steps, int modality, int B, cudaStream_t st, int step_begin, int step_end) {
  if (step_begin == 0) TRY(encoder(h, k, k.in_goal, k.in_state, modality, B, st));
  const size_t mrow = (size_t)h->Ld * 6 * h->d;
  for (int i = step_begin; i < step_end; ++i) {
    HeadArgs hd{};
    hd.x_in = k.x; hd.x_state = k.x; hd.x_aux = k.x2; hd.dbuf = k.dbuf; hd.sigmas = h->sigmas; hd.step = i; hd.n_steps = n_steps;
    switch (sampler) {
      case MDTB200_SAMPLER_DDIM: hd.mode = HEAD_DDIM; break;
      case MDTB200_SAMPLER_EULER: hd.mode = HEAD_EULER; break;
      case MDTB200_SAMPLER_HEUN: hd.mode = HEAD_HEUN1; break;
      case MDTB200_SAMPLER_DPMPP_2M: hd.mode = HEAD_DPMPP2M; break;
      case MDTB200_SAMPLER_EULER_ANCESTRAL: hd.mode = HEAD_EULER_ANC; hd.noise = k.noise + (size_t)i * h->noise_stride; hd.eta = h->anc_eta; break;
      default: return fail(h, MDTB200_EINVAL, "unknown sampler %d", sampler);
    }
    TRY(decoder_eval(h, k, k.x, h->mod + i * mrow, 0, h->sigmas + i, 0, 1, B, hd, st));
    if (sampler == MDTB200_SAMPLER_HEUN && i + 1 < n_steps) {   // 2nd-order correction; the last step (sigma_next = 0) is Euler
      HeadArgs h2 = hd;
      h2.mode = HEAD_HEUN2; h2.x_in = k.x2;
      TRY(decoder_eval(h, k, k.x2, h->mod + (i + 1) * mrow, 0, h->sigmas + i + 1, 0, 1, B, h2, st));
    }
  }
  return 0;
}


// ------------------------------------------------------------------------------------------ persistent fused decoder
// One evaluation of the score network inside a fused program: which sigma row of the AdaLN table it uses, what the output
// head does with the result and which action buffer the network reads.
struct EvalSpec { int sig; int mode; int step; const float* x_in; };

void free_plan(FusedPlan* pl) {
  if (!pl) return;
  if (pl->d_prog) cudaFree(pl->d_prog);
  if (pl->d_gemms) cudaFree(pl->d_gemms);
  if (pl->d_maps) cudaFree(pl->d_maps);
  delete pl;
}

void drop_graph(GraphEntry& ge) {
  if (ge.exec) cudaGraphExecDestroy(ge.exec);
  if (ge.exec2) cudaGraphExecDestroy(ge.exec2);
  free_plan(ge.plan);
  ge = GraphEntry{};
}

// Builds the phase program (fused_decoder.cuh) for `evals` consecutive score evaluations on the handle's static buffers.
//   mod / mod_stride     AdaLN table; row of evaluation e = mod + evals[e].sig * mrow (sampling: one sigma per step, shared by
//                        the batch, mod_stride = 0) or one row per sample (denoise API: evals[0].sig = 0, mod_stride = mrow)
//   sigma / sigma_stride the c_in scaling and the RAW / DENOISE head read sigma[evals[e].sig + sample * sigma_stride]
int build_fused_plan(MdtHandle* h, int B, const std::vector<EvalSpec>& evals, int n_steps, const float* mod, int mod_stride,
                     const float* sigma, int sigma_stride, int precondition, float* hout, FusedPlan** out) {
  const int d = h->d, T = h->T, Tc = h->Tc, Ld = h->Ld, C = h->fd_C, SPG = h->fd_SPG;
  const size_t Mp = ((size_t)h->cfg.max_batch * T + 127) / 128 * 128;
  const Weights& w = h->w;
  const Work k = h->work();
  const size_t mrow = (size_t)Ld * 6 * d;
  std::vector<CUtensorMap> maps(3 + 6 * Ld);
  const char* e;
  if ((e = h->tma.get(k.a16, (int)Mp, 2 * d, 2 * d, 128, &maps[0])) || (e = h->tma.get(k.y16, (int)Mp, 2 * d, 2 * d, 128, &maps[1])) ||
      (e = h->tma.get(k.h16, (int)Mp, 8 * d, 8 * d, 128, &maps[2])))
    return fail(h, MDTB200_ECUDA, "fused plan: %s", e);
  for (int l = 0; l < Ld; ++l) {
    const DecLayerW& L = w.dec[l];
    CUtensorMap* m = &maps[3 + 6 * l];
    if ((e = h->tma.get(L.wqkv16, 3 * d, 2 * d, 2 * d, 192, m + 0)) || (e = h->tma.get(L.wo16, d, 2 * d, 2 * d, 64, m + 1)) ||
        (e = h->tma.get(L.wq16, d, 2 * d, 2 * d, 64, m + 2)) || (e = h->tma.get(L.wco16, d, 2 * d, 2 * d, 64, m + 3)) ||
        (e = h->tma.get(L.wfc16, 4 * d, 2 * d, 2 * d, 256, m + 4)) || (e = h->tma.get(L.wproj16, d, 8 * d, 8 * d, d / 2, m + 5)))
      return fail(h, MDTB200_ECUDA, "fused plan: %s", e);
  }
  std::vector<fd::Phase> prog;
  std::vector<fd::GemmDesc> gemms;
  int acc = 0;
  auto blank = [&](int type) {
    fd::Phase ph;
    memset(&ph, 0, sizeof(ph));
    ph.type = type; ph.dep = (int)prog.size(); ph.head_mode = -1; ph.cta_mod = C;
    return ph;
  };
  // GEMM phase: every CTA computes W rows / output columns [cta * bn, +bn) from the full-K operand in tensor map a_map
  auto gemm_ph = [&](int a_map, int a_lo, int w_map, int nkb, int bn, int epi, fd::GemmDesc& g) {
    fd::Phase ph = blank(fd::PH_GEMM);
    ph.bn = bn; ph.epi = epi; ph.acc_buf = (acc++) & 1;
    ph.w_row_s1 = bn; ph.o_col_s1 = bn;
    memset(&g, 0, sizeof(g));
    g.a_map = a_map; g.a_lo_off = a_lo; g.w_map = w_map; g.nkb = nkb; g.bn = bn; g.acc_buf = ph.acc_buf; g.cta_mod = C;
    g.w_row_s1 = bn; g.w_lo_off = d;
    return ph;
  };
  auto push_gemm = [&](fd::Phase& ph, fd::GemmDesc& g) {
    g.p = (int)prog.size(); g.dep = ph.dep;
    prog.push_back(ph); gemms.push_back(g);
  };
  auto ln_fields = [&](fd::Phase& ph, const float* lw, const float* lb, const float* shift, const float* scale) {
    ph.ln = 1; ph.ln_w = lw; ph.ln_b = lb; ph.shift = shift; ph.scale = scale; ph.mod_stride = mod_stride;
    ph.out16 = k.a16; ph.ldo = 2 * d; ph.lo_off16 = d;
  };
  const float att_scale = 1.0f / sqrtf((float)h->hd);
  {  // action embedding of the first evaluation + LN1 / modulate of layer 0
    fd::Phase ph = blank(fd::PH_ROW);
    const float* m0 = mod + (size_t)evals[0].sig * mrow;
    ph.xh = k.xh; ph.embed = 1; ph.emb_x = evals[0].x_in; ph.ae_w = w.ae_w; ph.ae_b = w.ae_b;
    ph.emb_sigma = sigma + evals[0].sig; ph.emb_sigma_stride = sigma_stride; ph.precondition = precondition;
    ln_fields(ph, w.dec[0].ln1_w, w.dec[0].ln1_b, m0, m0 + d);
    prog.push_back(ph);
  }
  for (size_t ei = 0; ei < evals.size(); ++ei) {
    const EvalSpec& ev = evals[ei];
    const float* me = mod + (size_t)ev.sig * mrow;
    for (int l = 0; l < Ld; ++l) {
      const DecLayerW& L = w.dec[l];
      const float* ml = me + (size_t)l * 6 * d;
      const int mb = 3 + 6 * l;
      fd::GemmDesc g;
      {  // qkv = a . [Wq;Wk;Wv]^T + b
        fd::Phase ph = gemm_ph(0, d, mb + 0, d / 64, 192, fd::FE_STORE, g);
        ph.bias = L.bqkv; ph.out = k.qkv; ph.ldo = 3 * d;
        push_gemm(ph, g);
      }
      {  // causal self-attention -> y16
        fd::Phase ph = blank(fd::PH_ATTN);
        ph.q = k.qkv; ph.k = k.qkv + d; ph.v = k.qkv + 2 * d; ph.ldq = 3 * d; ph.ldkv = 3 * d; ph.Tq = T; ph.Tk = T; ph.causal = 1; ph.att_scale = att_scale;
        ph.out16 = k.y16; ph.ldo = 2 * d; ph.lo_off16 = d;
        prog.push_back(ph);
      }
      {  // x += gate_msa * (y . Wo^T + b)
        fd::Phase ph = gemm_ph(1, d, mb + 1, d / 64, 64, fd::FE_RESID, g);
        ph.bias = L.bo; ph.out = k.xh; ph.ldo = d; ph.gate = ml + 2 * d; ph.gate_stride = mod_stride;
        push_gemm(ph, g);
      }
      {  // LN3 -> a16
        fd::Phase ph = blank(fd::PH_ROW);
        ph.xh = k.xh;
        ln_fields(ph, L.ln3_w, L.ln3_b, nullptr, nullptr);
        prog.push_back(ph);
      }
      {  // q = a . Wq^T + b
        fd::Phase ph = gemm_ph(0, d, mb + 2, d / 64, 64, fd::FE_STORE, g);
        ph.bias = L.bq; ph.out = k.q; ph.ldo = d;
        push_gemm(ph, g);
      }
      {  // cross-attention over the cached context K/V (causal top-left mask)
        fd::Phase ph = blank(fd::PH_ATTN);
        ph.q = k.q; ph.ldq = d; ph.k = k.kv + (size_t)l * 2 * d; ph.v = k.kv + (size_t)l * 2 * d + d; ph.ldkv = Ld * 2 * d;
        ph.Tq = T; ph.Tk = Tc; ph.causal = 1; ph.att_scale = att_scale;
        ph.out16 = k.y16; ph.ldo = 2 * d; ph.lo_off16 = d;
        prog.push_back(ph);
      }
      {  // x += y . Wco^T + b
        fd::Phase ph = gemm_ph(1, d, mb + 3, d / 64, 64, fd::FE_RESID, g);
        ph.bias = L.bco; ph.out = k.xh; ph.ldo = d;
        push_gemm(ph, g);
      }
      {  // LN2 + modulate -> a16
        fd::Phase ph = blank(fd::PH_ROW);
        ph.xh = k.xh;
        ln_fields(ph, L.ln2_w, L.ln2_b, ml + 3 * d, ml + 4 * d);
        prog.push_back(ph);
      }
      {  // h16[:, slice] = split(gelu(a . Wfc[slice]^T + b))
        fd::Phase ph = gemm_ph(0, d, mb + 4, d / 64, 256, fd::FE_GELU16, g);
        ph.bias = L.bfc; ph.out16 = k.h16; ph.ldo = 8 * d; ph.lo_off16 = 4 * d;
        push_gemm(ph, g);
      }
      {  // c_proj as a (C/2 x 2) grid of (K slice of 512 hidden columns) x (N half): C/2 partial sums per output element
        fd::Phase ph = gemm_ph(2, 4 * d, mb + 5, 512 / 64, d / 2, fd::FE_PARTIAL, g);
        ph.cta_mod = C / 2; ph.w_row_s1 = 0; ph.w_row_s2 = d / 2; ph.o_col_s1 = 0; ph.o_col_s2 = d / 2;
        ph.out = h->fd_partial; ph.out_cta_stride = (long long)Mp * d; ph.ldo = d;
        g.cta_mod = C / 2; g.a_col_s1 = 512; g.w_row_s1 = 0; g.w_row_s2 = d / 2; g.w_col_s1 = 512; g.w_lo_off = 4 * d;
        push_gemm(ph, g);
      }
      {  // x += gate_mlp * (sum of partials + b); then the next consumer's LayerNorm (or the output head)
        fd::Phase ph = blank(fd::PH_ROW);
        ph.xh = k.xh; ph.n_part = C / 2; ph.part = h->fd_partial; ph.part_stride = (long long)Mp * d; ph.pbias = L.bproj;
        ph.rgate = ml + 5 * d; ph.rgate_stride = mod_stride;
        if (l + 1 < Ld) {
          const float* mn = me + (size_t)(l + 1) * 6 * d;
          ln_fields(ph, w.dec[l + 1].ln1_w, w.dec[l + 1].ln1_b, mn, mn + d);
        } else {
          ph.head_mode = ev.mode; ph.dln_w = w.dec_ln_w; ph.dln_b = w.dec_ln_b; ph.ap_w = w.ap_w; ph.ap_b = w.ap_b;
          ph.x_in = ev.x_in; ph.x_state = k.x; ph.x_aux = k.x2; ph.dbuf = k.dbuf; ph.hout = hout;
          ph.sigmas = h->sigmas; ph.step = ev.step; ph.n_steps = n_steps; ph.hsigma = sigma; ph.hsigma_stride = sigma_stride;
          if (ei + 1 < evals.size()) {
            const EvalSpec& nx = evals[ei + 1];
            const float* mn = mod + (size_t)nx.sig * mrow;
            ph.embed = 1; ph.emb_x = nullptr; ph.ae_w = w.ae_w; ph.ae_b = w.ae_b;
            ph.emb_sigma = sigma + nx.sig; ph.emb_sigma_stride = sigma_stride; ph.precondition = precondition;
            ln_fields(ph, w.dec[0].ln1_w, w.dec[0].ln1_b, mn, mn + d);
          }
        }
        prog.push_back(ph);
      }
    }
  }
  prog.push_back(blank(fd::PH_END));

  FusedPlan* pl = new (std::nothrow) FusedPlan();
  if (!pl) return fail(h, MDTB200_ENOMEM, "out of host memory");
  pl->n_phases = (int)prog.size();
  pl->n_rowgroups = (B + SPG - 1) / SPG;
  const int ng = pl->n_rowgroups < h->fd_groups_max ? pl->n_rowgroups : h->fd_groups_max;
  pl->grid = ng * C;
  if ((size_t)pl->n_rowgroups * 32 * sizeof(int) > h->fd_progress_bytes) { free_plan(pl); return fail(h, MDTB200_EINVAL, "fused plan: too many row groups"); }
  if (cudaMalloc(&pl->d_prog, prog.size() * sizeof(fd::Phase)) != cudaSuccess || cudaMalloc(&pl->d_maps, maps.size() * sizeof(CUtensorMap)) != cudaSuccess ||
      cudaMalloc(&pl->d_gemms, gemms.size() * sizeof(fd::GemmDesc)) != cudaSuccess ||
      cudaMemcpy(pl->d_gemms, gemms.data(), gemms.size() * sizeof(fd::GemmDesc), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(pl->d_prog, prog.data(), prog.size() * sizeof(fd::Phase), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(pl->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) != cudaSuccess) {
    free_plan(pl);
    cudaGetLastError();
    return fail(h, MDTB200_ENOMEM, "fused plan: device allocation failed");
  }
  fd::FusedParams& P = pl->params;
  P.prog = pl->d_prog; P.gemms = pl->d_gemms; P.n_gemms = (int)gemms.size(); P.maps = pl->d_maps; P.progress = h->fd_progress;
  P.B = B; P.T = T; P.A = h->A; P.d = d; P.H = h->H; P.hd = h->hd; P.C = C; P.SPG = SPG; P.n_rowgroups = pl->n_rowgroups;
  P.passes = h->cfg.precision == MDTB200_PREC_BF16X3 ? 3 : 1; P.sigma_data = h->cfg.sigma_data;
  P.trace = h->fd_trace;
  *out = pl;
  return 0;
}

int launch_fused(MdtHandle* h, const FusedPlan* pl, cudaStream_t st) {
  CUDA_TRY(h, cudaMemsetAsync(h->fd_progress, 0, (size_t)pl->n_rowgroups * 32 * sizeof(int), st));
  if (h->d == 384) fd::fused_decoder_kernel<3><<<pl->grid, fd::FD_THREADS, fd::SMEM_TOTAL, st>>>(pl->params);
  else fd::fused_decoder_kernel<4><<<pl->grid, fd::FD_THREADS, fd::SMEM_TOTAL, st>>>(pl->params);
  count_launch(h);
  return check_launch(h, "fused_decoder_kernel");
}

// evaluations of a sampler run (same sequence as sample_steps)
std::vector<EvalSpec> sampler_evals(const MdtHandle* h, int sampler, int n_steps) {
  std::vector<EvalSpec> ev;
  for (int i = 0; i < n_steps; ++i) {
    int mode = HEAD_DDIM;
    switch (sampler) {
      case MDTB200_SAMPLER_EULER: mode = HEAD_EULER; break;
      case MDTB200_SAMPLER_HEUN: mode = HEAD_HEUN1; break;
      case MDTB200_SAMPLER_DPMPP_2M: mode = HEAD_DPMPP2M; break;
      default: break;
    }
    ev.push_back(EvalSpec{i, mode, i, h->x});
    if (sampler == MDTB200_SAMPLER_HEUN && i + 1 < n_steps) ev.push_back(EvalSpec{i + 1, HEAD_HEUN2, i, h->x2});
  }
  return ev;
}

// The whole sampling call on the handle's static buffers (captured into a graph by mdtb200_sample).  The AdaLN
// table of all steps is computed once; then the batch is cut into `branches` independent sub-batches whose kernel
// chains run concurrently (fork/join through events -> parallel branches of the captured graph): every kernel of this
// path is latency- rather than throughput-bound at B=256, so concurrent chains fill the SMs the others leave idle.
//
// A call is captured as up to two graphs, [encoder + sampler steps 0..s0) and [steps s0..N): launching a ~2000-node graph costs
// ~0.35 ms of host time before the GPU starts, so the short first segment gets the GPU going and the long second one is
// launched while it runs (matters for the end-to-end path, where every call starts on an idle GPU).
int sample_body(MdtHandle* h, int sampler, int n_steps, int modality, int B, cudaStream_t st, int step_begin, int step_end) {
  if (step_begin == 0) TRY(sigma_path(h, h->sigmas, n_steps, st));   // one AdaLN row per step: sigma is shared by the batch
  const int nb = branch_count(h, B);
  h->cur_chains = nb < 1 ? 1 : nb;
  const Work base = h->work();
  const FusedPlan* plan = h->cur_plan;        // fused decoder: the branches only run the encoder, then ONE persistent kernel
  const int enc_end = plan ? 0 : step_end;
  int rc = 0;
  if (nb <= 1) {
    rc = sample_steps(h, base, sampler, n_steps, modality, B, st, step_begin, enc_end);
  } else {
    CUDA_TRY(h, cudaEventRecord(h->ev_fork, st));
    int mc_off = 0;
    for (int s = 0; s < nb && !rc; ++s) {
      const int b0 = (int)((long long)B * s / nb), b1 = (int)((long long)B * (s + 1) / nb);
      cudaStream_t bs = h->branch_streams[s];
      CUDA_TRY(h, cudaStreamWaitEvent(bs, h->ev_fork, 0));
      rc = sample_steps(h, h->slice(base, b0, mc_off), sampler, n_steps, modality, b1 - b0, bs, step_begin, enc_end);
      mc_off += round128((b1 - b0) * h->Tc);
      CUDA_TRY(h, cudaEventRecord(h->ev_join[s], bs));
      CUDA_TRY(h, cudaStreamWaitEvent(st, h->ev_join[s], 0));
    }
  }
  if (rc || !plan) return rc;
  return launch_fused(h, plan, st);
}

// ------------------------------------------------------------------------------------------ weights

const float* need(MdtHandle* h, const std::string& name, int64_t numel, bool optional, bool& ok) {
  auto it = h->bound.find(name);
  if (it == h->bound.end()) {
    if (!optional) { ok = false; fail(h, MDTB200_ESTATE, "weight '%s' was not bound", name.c_str()); }
    return nullptr;
  }
  if (it->second.numel != numel) {
    ok = false;
    fail(h, MDTB200_EINVAL, "weight '%s': expected %lld elements, got %lld", name.c_str(), (long long)numel, (long long)it->second.numel);
    return nullptr;
  }
  return it->second.ptr;
}

// copies `numel` floats from a bound tensor into the arena, returns the arena pointer
struct Packer {
  MdtHandle* h; cudaStream_t st; bool ok = true; std::string p;
  float* take(size_t n) {
    if (h->arena_used + n > h->arena_floats) { ok = false; fail(h, MDTB200_ENOMEM, "weight arena overflow"); return nullptr; }
    float* r = h->arena + h->arena_used;
    h->arena_used += (n + 31) / 32 * 32;   // keep 128-byte alignment
    return r;
  }
  const float* copy(const std::string& name, int64_t numel, bool optional = false) {
    const float* src = need(h, p + name, numel, optional, ok);
    if (!src || !ok) return nullptr;
    float* dst = take((size_t)numel);
    if (!dst) return nullptr;
    if (cudaMemcpyAsync(dst, src, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { ok = false; fail(h, MDTB200_ECUDA, "memcpy of %s failed", name.c_str()); }
    return dst;
  }
  // concatenation of several tensors (all required)
  const float* concat(const std::vector<std::pair<std::string, int64_t>>& parts) {
    size_t total = 0;
    for (auto& q : parts) total += (size_t)q.second;
    float* dst = take(total);
    if (!dst) return nullptr;
    size_t off = 0;
    for (auto& q : parts) {
      const float* src = need(h, p + q.first, q.second, false, ok);
      if (!ok) return nullptr;
      if (cudaMemcpyAsync(dst + off, src, (size_t)q.second * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) { ok = false; fail(h, MDTB200_ECUDA, "memcpy of %s failed", q.first.c_str()); return nullptr; }
      off += (size_t)q.second;
    }
    return dst;
  }
  // split-bf16 copy (rows, 2*cols) of an arena matrix
  const __nv_bfloat16* split(const float* src, int64_t rows, int cols) {
    if (!src || !ok || !h->arena16) return nullptr;
    size_t n = (size_t)rows * cols * 2;
    if (h->arena16_used + n > h->arena16_elems) { ok = false; fail(h, MDTB200_ENOMEM, "bf16 weight arena overflow"); return nullptr; }
    __nv_bfloat16* dst = h->arena16 + h->arena16_used;
    h->arena16_used += (n + 63) / 64 * 64;
    int64_t total = rows * cols;
    split_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, dst, rows, cols);
    return dst;
  }
};

size_t arena_size(const MdtHandle* h) {
  const size_t d = h->d, G = h->cfg.goal_dim, O = h->cfg.obs_dim, A = h->A;
  size_t n = 0;
  n += 2 * (2 * d * G + 2 * d + d * 2 * d + d);            // goal_emb + lang_emb
  n += 2 * (d * O + d) + 64 * d;                           // tok_emb, incam_embed, pos_emb
  n += h->Le * (12 * d * d + 3 * d + 8 * d + 64 * 12);      // encoder layers (+optional biases/padding)
  n += h->Ld * (14 * d * d + 3 * d + d + 12 * d + 64 * 18); // decoder layers
  n += h->Ld * (2 * d * d + 2 * d) + h->Ld * (6 * d * d + 6 * d) + h->Ld * (d + 64);
  n += 4 * d + 2 * d * d * 2 + 3 * d + A * d * 2 + d + A + 64 * 64;
  return n + (1 << 16);
}

int pack_weights(MdtHandle* h, cudaStream_t st) {
  const int64_t d = h->d, G = h->cfg.goal_dim, O = h->cfg.obs_dim, A = h->A;
  h->arena_used = 0; h->arena16_used = 0; h->last_code = 0;
  Packer pk{h, st};
  pk.p = "inner_model.";
  Weights& w = h->w;
  w = Weights{};
  w.goal0_w = pk.copy("goal_emb.0.weight", 2 * d * G); w.goal0_b = pk.copy("goal_emb.0.bias", 2 * d);
  w.goal2_w = pk.copy("goal_emb.2.weight", d * 2 * d); w.goal2_b = pk.copy("goal_emb.2.bias", d);
  w.lang0_w = pk.copy("lang_emb.0.weight", 2 * d * G, true);
  if (w.lang0_w) {
    w.lang0_b = pk.copy("lang_emb.0.bias", 2 * d); w.lang2_w = pk.copy("lang_emb.2.weight", d * 2 * d); w.lang2_b = pk.copy("lang_emb.2.bias", d);
  }
  w.tok_w = pk.copy("tok_emb.weight", d * O); w.tok_b = pk.copy("tok_emb.bias", d);
  w.goal0_w16 = pk.split(w.goal0_w, 2 * d, (int)G); w.goal2_w16 = pk.split(w.goal2_w, d, (int)(2 * d));
  if (w.lang0_w) { w.lang0_w16 = pk.split(w.lang0_w, 2 * d, (int)G); w.lang2_w16 = pk.split(w.lang2_w, d, (int)(2 * d)); }
  w.tok_w16 = pk.split(w.tok_w, d, (int)O);
  if (h->cfg.variant == MDTB200_VARIANT_MDT) {
    w.incam_w = pk.copy("incam_embed.weight", d * O); w.incam_b = pk.copy("incam_embed.bias", d);
    w.incam_w16 = pk.split(w.incam_w, d, (int)O);
    // pos_emb (1, goal_seq_len + action_seq_len, d): rows 0 and 1 are used by the encoder
    auto it = h->bound.find("inner_model.pos_emb");
    if (it == h->bound.end() || it->second.numel < 2 * d) return fail(h, MDTB200_ESTATE, "weight 'inner_model.pos_emb' missing or too small");
    float* dst = pk.take((size_t)2 * d);
    if (dst) cudaMemcpyAsync(dst, it->second.ptr, (size_t)2 * d * 4, cudaMemcpyDeviceToDevice, st);
    w.pos_emb = dst;
  }
  w.enc.resize(h->Le);
  for (int l = 0; l < h->Le && pk.ok; ++l) {
    EncLayerW& L = w.enc[l];
    std::string b = "encoder.blocks." + std::to_string(l) + ".";
    L.ln1_w = pk.copy(b + "ln_1.weight", d); L.ln1_b = pk.copy(b + "ln_1.bias", d, true);
    L.wqkv = pk.concat({{b + "attn.query.weight", d * d}, {b + "attn.key.weight", d * d}, {b + "attn.value.weight", d * d}});
    L.bqkv = pk.concat({{b + "attn.query.bias", d}, {b + "attn.key.bias", d}, {b + "attn.value.bias", d}});
    L.wo = pk.copy(b + "attn.c_proj.weight", d * d); L.bo = pk.copy(b + "attn.c_proj.bias", d, true);
    L.ln2_w = pk.copy(b + "ln_2.weight", d); L.ln2_b = pk.copy(b + "ln_2.bias", d, true);
    L.wfc = pk.copy(b + "mlp.c_fc.weight", 4 * d * d); L.bfc = pk.copy(b + "mlp.c_fc.bias", 4 * d, true);
    L.wproj = pk.copy(b + "mlp.c_proj.weight", 4 * d * d); L.bproj = pk.copy(b + "mlp.c_proj.bias", d, true);
    L.wqkv16 = pk.split(L.wqkv, 3 * d, (int)d); L.wo16 = pk.split(L.wo, d, (int)d);
    L.wfc16 = pk.split(L.wfc, 4 * d, (int)d); L.wproj16 = pk.split(L.wproj, d, (int)(4 * d));
  }
  w.enc_ln_w = pk.copy("encoder.ln.weight", d); w.enc_ln_b = pk.copy("encoder.ln.bias", d, true);
  w.dec.resize(h->Ld);
  std::vector<std::pair<std::string, int64_t>> kvw, kvb, modw, modb, bqv;
  for (int l = 0; l < h->Ld && pk.ok; ++l) {
    DecLayerW& L = w.dec[l];
    std::string b = "decoder.blocks." + std::to_string(l) + ".";
    L.ln1_w = pk.copy(b + "ln_1.weight", d); L.ln1_b = pk.copy(b + "ln_1.bias", d, true);
    L.wqkv = pk.concat({{b + "attn.query.weight", d * d}, {b + "attn.key.weight", d * d}, {b + "attn.value.weight", d * d}});
    L.bqkv = pk.concat({{b + "attn.query.bias", d}, {b + "attn.key.bias", d}, {b + "attn.value.bias", d}});
    L.wo = pk.copy(b + "attn.c_proj.weight", d * d); L.bo = pk.copy(b + "attn.c_proj.bias", d, true);
    L.ln3_w = pk.copy(b + "ln3.weight", d); L.ln3_b = pk.copy(b + "ln3.bias", d);
    L.wq = pk.copy(b + "cross_att.query.weight", d * d); L.bq = pk.copy(b + "cross_att.query.bias", d);
    L.wco = pk.copy(b + "cross_att.c_proj.weight", d * d); L.bco = pk.copy(b + "cross_att.c_proj.bias", d, true);
    L.ln2_w = pk.copy(b + "ln_2.weight", d); L.ln2_b = pk.copy(b + "ln_2.bias", d, true);
    L.wfc = pk.copy(b + "mlp.c_fc.weight", 4 * d * d); L.bfc = pk.copy(b + "mlp.c_fc.bias", 4 * d, true);
    L.wproj = pk.copy(b + "mlp.c_proj.weight", 4 * d * d); L.bproj = pk.copy(b + "mlp.c_proj.bias", d, true);
    L.wqkv16 = pk.split(L.wqkv, 3 * d, (int)d); L.wo16 = pk.split(L.wo, d, (int)d); L.wq16 = pk.split(L.wq, d, (int)d);
    L.wco16 = pk.split(L.wco, d, (int)d); L.wfc16 = pk.split(L.wfc, 4 * d, (int)d); L.wproj16 = pk.split(L.wproj, d, (int)(4 * d));
    if (h->cross_fused && L.wq && L.wco && pk.ok) {
      const size_t n = (size_t)h->H * d * 128;
      if (h->arena16_used + 2 * n > h->arena16_elems) { pk.ok = false; fail(h, MDTB200_ENOMEM, "bf16 weight arena overflow"); }
      else {
        __nv_bfloat16* wg = h->arena16 + h->arena16_used; __nv_bfloat16* wu = wg + n;
        h->arena16_used += 2 * n;
        const int tot = h->H * (int)d * 64;
        pack_cross_weights_kernel<<<(tot + 255) / 256, 256, 0, st>>>(L.wq, L.wco, wg, wu, (int)d, h->H, h->hd);
        L.wgx16 = wg; L.wux16 = wu;
      }
    }
    bqv.push_back({b + "cross_att.query.bias", d});
    kvw.push_back({b + "cross_att.key.weight", d * d}); kvw.push_back({b + "cross_att.value.weight", d * d});
    kvb.push_back({b + "cross_att.key.bias", d}); kvb.push_back({b + "cross_att.value.bias", d});
    modw.push_back({b + "adaLN_zero.modulation.1.weight", 6 * d * d}); modb.push_back({b + "adaLN_zero.modulation.1.bias", 6 * d});
  }
  if (pk.ok) {
    w.wkv_all = pk.concat(kvw); w.bkv_all = pk.concat(kvb);
    w.wmod_all = pk.concat(modw); w.bmod_all = pk.concat(modb); w.bq_all = pk.concat(bqv);
    w.wkv_all16 = pk.split(w.wkv_all, (int64_t)h->Ld * 2 * d, (int)d);
  }
  w.dec_ln_w = pk.copy("decoder.ln.weight", d); w.dec_ln_b = pk.copy("decoder.ln.bias", d, true);
  w.sig1_w = pk.copy("sigma_emb.1.weight", 2 * d * d); w.sig1_b = pk.copy("sigma_emb.1.bias", 2 * d);
  w.sig3_w = pk.copy("sigma_emb.3.weight", d * 2 * d); w.sig3_b = pk.copy("sigma_emb.3.bias", d);
  w.ae_w = pk.copy("action_emb.weight", d * A); w.ae_b = pk.copy("action_emb.bias", d);
  w.ap_w = pk.copy("action_pred.weight", A * d); w.ap_b = pk.copy("action_pred.bias", A);
  if (!pk.ok) return h->last_code ? h->last_code : MDTB200_ESTATE;
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

template <typename T>
int dev_alloc(MdtHandle* h, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, n * sizeof(T));
  if (e != cudaSuccess) return fail(h, MDTB200_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int check_ready(MdtHandle* h, int B) {
  if (!h) return MDTB200_EINVAL;
  if (!h->committed) return fail(h, MDTB200_ESTATE, "weights not committed (call mdtb200_commit_weights)");
  if (B < 1 || B > h->cfg.max_batch) return fail(h, MDTB200_EINVAL, "batch %d outside [1, max_batch=%d]", B, h->cfg.max_batch);
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != h->device) return fail(h, MDTB200_ESTATE, "handle belongs to device %d but device %d is current", h->device, dev);
  return 0;
}

constexpr int MAX_STEPS = 256;

}  // namespace

// =========================================================================================== C ABI

extern "C" {

MDTB200_API int mdtb200_abi_version(void) { return MDTB200_ABI_VERSION; }

MDTB200_API const char* mdtb200_last_error(const MdtHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

MDTB200_API int mdtb200_create(const MdtConfig* cfg, MdtHandle** out) {
  if (!cfg || !out) return fail(nullptr, MDTB200_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MDTB200_ABI_VERSION) return fail(nullptr, MDTB200_EINVAL, "abi_version %d != %d", cfg->abi_version, MDTB200_ABI_VERSION);
  const int d = cfg->embed_dim;
  if (d < 128 || d % 128 != 0 || d > 1024 || d / 128 == 5 || d / 128 == 7) return fail(nullptr, MDTB200_EUNSUPPORTED, "embed_dim %d unsupported (128,256,384,512,768,1024)", d);
  if (cfg->n_heads < 1 || d % cfg->n_heads != 0 || d / cfg->n_heads > ATT_MAXHD) return fail(nullptr, MDTB200_EUNSUPPORTED, "n_heads %d unsupported for d=%d (head_dim <= %d)", cfg->n_heads, d, ATT_MAXHD);
  if (cfg->action_seq_len < 1 || cfg->action_seq_len > ATT_MAXT) return fail(nullptr, MDTB200_EUNSUPPORTED, "action_seq_len %d > %d", cfg->action_seq_len, ATT_MAXT);
  if (cfg->n_state_tokens < 1 || 1 + cfg->n_state_tokens > ATT_MAXT) return fail(nullptr, MDTB200_EUNSUPPORTED, "n_state_tokens %d unsupported", cfg->n_state_tokens);
  if (cfg->variant != MDTB200_VARIANT_MDTV && cfg->variant != MDTB200_VARIANT_MDT) return fail(nullptr, MDTB200_EINVAL, "unknown variant %d", cfg->variant);
  if (cfg->variant == MDTB200_VARIANT_MDT && cfg->n_state_tokens != 2) return fail(nullptr, MDTB200_EINVAL, "MDT variant has exactly 2 state tokens");
  if (cfg->goal_dim % 16 || cfg->obs_dim % 16) return fail(nullptr, MDTB200_EUNSUPPORTED, "goal_dim/obs_dim must be multiples of 16");
  if (cfg->action_dim < 1 || cfg->action_dim > 32) return fail(nullptr, MDTB200_EINVAL, "action_dim %d", cfg->action_dim);
  if (cfg->n_enc_layers < 0 || cfg->n_dec_layers < 1 || cfg->max_batch < 1) return fail(nullptr, MDTB200_EINVAL, "bad layer count / max_batch");
  if (cfg->precision < MDTB200_PREC_FP32 || cfg->precision > MDTB200_PREC_BF16) return fail(nullptr, MDTB200_EINVAL, "unknown precision %d", cfg->precision);
  if (!(cfg->sigma_data > 0.f)) return fail(nullptr, MDTB200_EINVAL, "sigma_data must be > 0");

  MdtHandle* h = new (std::nothrow) MdtHandle();
  if (!h) return fail(nullptr, MDTB200_ENOMEM, "out of host memory");
  h->cfg = *cfg;
  h->d = d; h->H = cfg->n_heads; h->hd = d / cfg->n_heads; h->T = cfg->action_seq_len; h->Ts = cfg->n_state_tokens; h->Tc = 1 + cfg->n_state_tokens;
  h->A = cfg->action_dim; h->Le = cfg->n_enc_layers; h->Ld = cfg->n_dec_layers;
  int rc = 0;
  auto bail = [&](int code) { g_create_error = h->err; mdtb200_destroy(h); return code; };
  if (cudaGetDevice(&h->device) != cudaSuccess) { fail(h, MDTB200_ECUDA, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); return bail(MDTB200_ECUDA); }
  if (cfg->precision != MDTB200_PREC_FP32) {
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, h->device);
    if (prop.major != 10) { fail(h, MDTB200_EUNSUPPORTED, "tensor-core precision needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor); return bail(MDTB200_EUNSUPPORTED); }
    const char* e = h->tma.init();
    if (e) { fail(h, MDTB200_ECUDA, "TMA descriptor encoder unavailable: %s", e); return bail(MDTB200_ECUDA); }
    e = tc::configure_kernels();
    if (e) { fail(h, MDTB200_ECUDA, "tcgen05 kernel configuration failed: %s", e); return bail(MDTB200_ECUDA); }
  }
  const size_t B = cfg->max_batch, M = B * h->T, Dd = d;
  h->arena_floats = arena_size(h);
  if ((rc = dev_alloc(h, &h->arena, h->arena_floats))) return bail(rc);
  if (cfg->precision != MDTB200_PREC_FP32) {
    h->arena16_elems = 2 * ((size_t)h->Le * 12 * Dd * Dd + (size_t)h->Ld * 16 * Dd * Dd + 2 * (2 * Dd * cfg->goal_dim + 2 * Dd * Dd) + 2 * Dd * cfg->obs_dim) + (1 << 16) +
                       (size_t)h->Ld * 2 * h->H * Dd * 128;
    if ((rc = dev_alloc(h, &h->arena16, h->arena16_elems))) return bail(rc);
  }
  h->mod_rows = (int)(B > (size_t)MAX_STEPS ? B : (size_t)MAX_STEPS);
  const size_t R = h->mod_rows;
  if ((rc = dev_alloc(h, &h->in_goal, B * cfg->goal_dim)) || (rc = dev_alloc(h, &h->in_state, B * h->Ts * cfg->obs_dim)) ||
      (rc = dev_alloc(h, &h->x, M * h->A)) || (rc = dev_alloc(h, &h->x2, M * h->A)) || (rc = dev_alloc(h, &h->dbuf, M * h->A)) ||
      (rc = dev_alloc(h, &h->sigmas, (size_t)MAX_STEPS + 8 + B)) ||
      (rc = dev_alloc(h, &h->gh, B * 2 * Dd)) || (rc = dev_alloc(h, &h->xe, M * Dd)) || (rc = dev_alloc(h, &h->ctx, M * Dd)) ||
      (rc = dev_alloc(h, &h->kv, B * h->Tc * h->Ld * 2 * Dd)) || (rc = dev_alloc(h, &h->xh, M * Dd)) || (rc = dev_alloc(h, &h->a, M * Dd)) ||
      (rc = dev_alloc(h, &h->qkv, M * 3 * Dd)) || (rc = dev_alloc(h, &h->y, M * Dd)) || (rc = dev_alloc(h, &h->hbuf, M * 4 * Dd)) || (rc = dev_alloc(h, &h->q, M * Dd)) ||
      (rc = dev_alloc(h, &h->pe, R * Dd)) || (rc = dev_alloc(h, &h->sh, R * 2 * Dd)) || (rc = dev_alloc(h, &h->cs, R * Dd)) || (rc = dev_alloc(h, &h->mod, R * h->Ld * 6 * Dd)))
    return bail(rc);
  if (cfg->precision != MDTB200_PREC_FP32) {
    // row counts padded to the 128-row MMA tile so TMA boxes never leave the allocation
    const size_t Mp = (M + 127) / 128 * 128;
    if ((rc = dev_alloc(h, &h->a16, Mp * 2 * Dd)) || (rc = dev_alloc(h, &h->y16, Mp * 2 * Dd)) || (rc = dev_alloc(h, &h->h16, Mp * 8 * Dd))) return bail(rc);
    cudaMemset(h->a16, 0, Mp * 2 * Dd * 2); cudaMemset(h->y16, 0, Mp * 2 * Dd * 2); cudaMemset(h->h16, 0, Mp * 8 * Dd * 2);
  }
  {
    // algebraic cross-attention (cross_row_kernel): needs the shapes its register / thread budget was written for
    const char* off = getenv("MDTB200_CROSS_FUSED");
    const int HT = h->H * h->Tc;
    const size_t smem = cross_row_smem_bytes(d, h->T, h->Tc, h->H);
    if (cfg->precision != MDTB200_PREC_FP32 && !(off && off[0] == '0') && (d == 384 || d == 512) && h->hd <= 64 && HT * (d / 4) <= 8 * CR_THREADS &&
        h->T * 32 < CR_THREADS && h->T * HT <= CR_THREADS * 4 && h->T * h->H <= CR_THREADS && smem <= 200 * 1024) {
      if (cudaFuncSetAttribute(cross_row_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
          cudaFuncSetAttribute(cross_row_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        fail(h, MDTB200_ECUDA, "cross_row_kernel needs %zu bytes of shared memory", smem); return bail(MDTB200_ECUDA);
      }
      h->cross_rows = (size_t)h->H * (round128((int)B * h->Tc) + 128 * MdtHandle::MAX_BRANCHES);
      h->ka_layer_stride = h->cross_rows * 128; h->tab_layer_stride = h->cross_rows * Dd; h->ctab_layer_stride = B * HT;
      if ((rc = dev_alloc(h, &h->ka16, h->Ld * h->ka_layer_stride)) || (rc = dev_alloc(h, &h->va16, h->Ld * h->ka_layer_stride)) ||
          (rc = dev_alloc(h, &h->gtab, h->Ld * h->tab_layer_stride)) || (rc = dev_alloc(h, &h->utab, h->Ld * h->tab_layer_stride)) ||
          (rc = dev_alloc(h, &h->ctab, h->Ld * h->ctab_layer_stride))) return bail(rc);
      cudaMemset(h->ka16, 0, h->Ld * h->ka_layer_stride * 2); cudaMemset(h->va16, 0, h->Ld * h->ka_layer_stride * 2);
      h->cross_fused = true;
    }
  }
  if (cfg->precision != MDTB200_PREC_FP32 && (d == 384 || d == 512) && 128 / h->T >= 1 && h->A <= 32) {
    // persistent fused decoder: C = d / 64 CTAs per row group of SPG = 128 / T samples (fused_decoder.cuh)
    const char* on = getenv("MDTB200_FUSED");     // opt-in: measured slower than the multi-branch graph (profiles/r02_fused_decoder.md)
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, h->device);
    h->fd_C = d / 64; h->fd_SPG = 128 / h->T; h->fd_groups_max = prop.multiProcessorCount / h->fd_C;
    const size_t att = (size_t)2 * (h->T + 2 * (h->T > h->Tc ? h->T : h->Tc)) * (d + 4) * 4 + (size_t)2 * h->H * h->T * (h->T + 1) * 4;
    if (on && on[0] == '1' && h->fd_groups_max >= 1 && att <= (size_t)fd::ROW_SCR_BYTES) {
      const char* e = fd::configure_fused();
      if (e) { fail(h, MDTB200_ECUDA, "fused decoder configuration failed: %s", e); return bail(MDTB200_ECUDA); }
      const size_t Mp = (M + 127) / 128 * 128;
      h->fd_progress_bytes = ((B + h->fd_SPG - 1) / h->fd_SPG) * 32 * sizeof(int);
      if ((rc = dev_alloc(h, &h->fd_partial, (size_t)(h->fd_C / 2) * Mp * Dd)) || (rc = dev_alloc(h, &h->fd_progress, h->fd_progress_bytes / sizeof(int)))) return bail(rc);
      h->fused = true;
      if (getenv("MDTB200_FUSED_TRACE")) {
        if ((rc = dev_alloc(h, &h->fd_trace, (size_t)fd::TRACE_PHASES * 160 * 8))) return bail(rc);
        cudaMemset(h->fd_trace, 0, (size_t)fd::TRACE_PHASES * 160 * 8 * 8);
      }
    }
  }
  {
    size_t att = attention_smem_bytes(d, h->H, h->T, h->T);
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att) != cudaSuccess) {
      fail(h, MDTB200_ECUDA, "attention kernel needs %zu bytes of shared memory", att); return bail(MDTB200_ECUDA);
    }
  }
  if (cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) { fail(h, MDTB200_ECUDA, "cudaStreamCreate failed"); return bail(MDTB200_ECUDA); }
  for (int i = 0; i < MdtHandle::MAX_BRANCHES; ++i) {
    if (cudaStreamCreateWithFlags(&h->branch_streams[i], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join[i], cudaEventDisableTiming) != cudaSuccess) { fail(h, MDTB200_ECUDA, "stream/event creation failed"); return bail(MDTB200_ECUDA); }
  }
  if (cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess) { fail(h, MDTB200_ECUDA, "event creation failed"); return bail(MDTB200_ECUDA); }
  if (const char* e = getenv("MDTB200_BRANCHES")) h->branches = atoi(e);
  if (const char* e = getenv("MDTB200_PDL_LATE")) h->pdl_late = atoi(e) != 0;
  if (cfg->precision == MDTB200_PREC_BF16X3) {
    if (const char* e = getenv("MDTB200_CPROJ_SPLITS")) h->cproj_splits = atoi(e);
    if (h->cproj_splits < 0 || h->cproj_splits > MdtHandle::SK_MAX_SPLITS) h->cproj_splits = 0;
    if (h->cproj_splits != 1) {
      const size_t tiles = ((size_t)cfg->max_batch * h->T + 127) / 128 + (size_t)cfg->max_batch / 32 + 2;
      int rc2 = 0;
      if ((rc2 = dev_alloc(h, &h->sk_ws, tiles * MdtHandle::SK_MAX_SPLITS * 128 * h->d)) || (rc2 = dev_alloc(h, &h->sk_cnt, tiles * (h->d / 64)))) return bail(rc2);
      cudaMemset(h->sk_cnt, 0, tiles * (h->d / 64) * sizeof(unsigned));
    }
  } else {
    h->cproj_splits = 1;
  }
  *out = h;
  return 0;
}

MDTB200_API void mdtb200_destroy(MdtHandle* h) {
  if (!h) return;
  for (auto& kv : h->graphs) drop_graph(kv.second);
  free_plan(h->denoise_plan);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (int i = 0; i < MdtHandle::MAX_BRANCHES; ++i) {
    if (h->branch_streams[i]) cudaStreamDestroy(h->branch_streams[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

MDTB200_API int mdtb200_bind_weight(MdtHandle* h, const char* name, const float* dev_ptr, int64_t numel) {
  if (!h) return MDTB200_EINVAL;
  if (!name || !dev_ptr || numel <= 0) return fail(h, MDTB200_EINVAL, "bind_weight: bad argument");
  h->bound[name] = Bound{dev_ptr, numel};
  return 0;
}

MDTB200_API int mdtb200_commit_weights(MdtHandle* h, void* stream) {
  if (!h) return MDTB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  h->committed = false;
  h->ctx_B = 0;
  // cached graphs reference arena addresses: drop them, the layout may change with the set of bound tensors
  for (auto& kv : h->graphs) drop_graph(kv.second);
  h->graphs.clear();
  free_plan(h->denoise_plan); h->denoise_plan = nullptr; h->denoise_plan_B = 0;
  h->tma.cache.clear();
  int rc = pack_weights(h, st);
  h->bound.clear();   // source pointers are not retained
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(st));
  h->committed = true;
  return 0;
}

MDTB200_API int mdtb200_encode(MdtHandle* h, const float* goal, const float* state, int modality, int B, float* ctx_out, void* stream) {
  TRY(check_ready(h, B));
  if (!goal || !state) return fail(h, MDTB200_EINVAL, "encode: null input");
  cudaStream_t st = (cudaStream_t)stream;
  TRY(encoder(h, h->work(), goal, state, modality, B, st));
  if (ctx_out) CUDA_TRY(h, cudaMemcpyAsync(ctx_out, h->ctx, (size_t)B * h->Tc * h->d * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

MDTB200_API int mdtb200_set_context(MdtHandle* h, const float* ctx, int B, void* stream) {
  TRY(check_ready(h, B));
  if (!ctx) return fail(h, MDTB200_EINVAL, "set_context: null input");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(h, cudaMemcpyAsync(h->ctx, ctx, (size_t)B * h->Tc * h->d * 4, cudaMemcpyDeviceToDevice, st));
  const bool tcp = use_tc(h);
  if (tcp) {   // same split-bf16 operand the encoder's final LayerNorm would have produced
    int64_t n = (int64_t)B * h->Tc * h->d;
    split_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->ctx, h->a16, (int64_t)B * h->Tc, h->d);
    count_launch(h);
    TRY(check_launch(h, "split_weights_kernel"));
  }
  TRY(compute_kv(h, h->work(), B, st, tcp));
  h->ctx_B = B;
  return 0;
}

MDTB200_API int mdtb200_denoise(MdtHandle* h, const float* x, const float* sigma, int B, int precondition, float* out, void* stream) {
  TRY(check_ready(h, B));
  if (!x || !sigma || !out) return fail(h, MDTB200_EINVAL, "denoise: null argument");
  if (h->ctx_B != B) return fail(h, MDTB200_ESTATE, "denoise: cached context is for batch %d, call has batch %d (encode first)", h->ctx_B, B);
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_chains = 1;
  TRY(sigma_path(h, sigma, B, st));
  HeadArgs hd{};
  hd.mode = precondition ? HEAD_DENOISE : HEAD_RAW; hd.x_in = x; hd.out = out; hd.sigma = sigma; hd.sigma_stride = 1;
  return decoder_eval(h, h->work(), x, h->mod, h->Ld * 6 * h->d, sigma, 1, precondition, B, hd, st);
}

static int capture_segment(MdtHandle* h, int sampler, int n_steps, int modality, int B, int step_begin, int step_end, cudaGraphExec_t* exec) {
  cudaGraph_t graph = nullptr;
  CUDA_TRY(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
  h->capturing = true;
  int rc = sample_body(h, sampler, n_steps, modality, B, h->cap_stream, step_begin, step_end);
  h->capturing = false;
  cudaError_t e = cudaStreamEndCapture(h->cap_stream, &graph);
  if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
  if (e != cudaSuccess) return fail(h, MDTB200_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(h, MDTB200_ECUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
  return 0;
}

static int get_graph(MdtHandle* h, int sampler, int n_steps, int modality, int B, GraphEntry** out) {
  GraphKey key{B, n_steps, sampler, modality};
  auto it = h->graphs.find(key);
  if (it != h->graphs.end()) { it->second.last_use = ++h->use_clock; *out = &it->second; return 0; }
  // bounded cache (a ~2000-node executable graph per distinct (B, N, sampler, modality)): evict the least recently used
  while (h->graphs.size() >= MdtHandle::MAX_GRAPHS) {
    auto lru = h->graphs.begin();
    for (auto jt = h->graphs.begin(); jt != h->graphs.end(); ++jt) if (jt->second.last_use < lru->second.last_use) lru = jt;
    drop_graph(lru->second);
    h->graphs.erase(lru);
  }
  GraphEntry ge{};
  h->capture_count = 0;
  if (h->fused && sampler != MDTB200_SAMPLER_EULER_ANCESTRAL) TRY(build_fused_plan(h, B, sampler_evals(h, sampler, n_steps), n_steps, h->mod, 0, h->sigmas, 0, 1, nullptr, &ge.plan));
  static const bool split = !getenv("MDTB200_SINGLE_GRAPH");
  const int s0 = (!ge.plan && split && n_steps > 3) ? 2 : n_steps;
  h->cur_plan = ge.plan;
  int rc = capture_segment(h, sampler, n_steps, modality, B, 0, s0, &ge.exec);
  h->cur_plan = nullptr;
  if (rc) { free_plan(ge.plan); return rc; }
  if (s0 < n_steps) {
    rc = capture_segment(h, sampler, n_steps, modality, B, s0, n_steps, &ge.exec2);
    if (rc) { drop_graph(ge); return rc; }
  }
  ge.kernels = h->capture_count;
  ge.last_use = ++h->use_clock;
  auto ins = h->graphs.emplace(key, ge);
  *out = &ins.first->second;
  return 0;
}

static int check_sample_args(MdtHandle* h, int sampler, int n_steps, int B, const void* a, const void* b, const void* c, const void* d) {
  TRY(check_ready(h, B));
  if (!a || !b || !c || !d) return fail(h, MDTB200_EINVAL, "sample: null argument");
  if (n_steps < 1 || n_steps > MAX_STEPS) return fail(h, MDTB200_EINVAL, "n_steps %d outside [1, %d]", n_steps, MAX_STEPS);
  if (sampler < MDTB200_SAMPLER_DDIM || sampler > MDTB200_SAMPLER_EULER_ANCESTRAL) return fail(h, MDTB200_EINVAL, "unknown sampler %d", sampler);
  return 0;
}

static int sample_impl(MdtHandle* h, int sampler, const float* sigmas, int n_steps, const float* goal, const float* state, int modality, int B,
                       float* x_inout, cudaStream_t st, cudaMemcpyKind in_kind, cudaMemcpyKind out_kind) {
  TRY(check_sample_args(h, sampler, n_steps, B, sigmas, goal, state, x_inout));
  GraphEntry* ge = nullptr;
  TRY(get_graph(h, sampler, n_steps, modality != 0, B, &ge));
  const size_t xb = (size_t)B * h->T * h->A * 4;
  CUDA_TRY(h, cudaMemcpyAsync(h->sigmas, sigmas, (size_t)(n_steps + 1) * 4, in_kind, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->in_goal, goal, (size_t)B * h->cfg.goal_dim * 4, in_kind, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->in_state, state, (size_t)B * h->Ts * h->cfg.obs_dim * 4, in_kind, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->x, x_inout, xb, in_kind, st));
  CUDA_TRY(h, cudaGraphLaunch(ge->exec, st));
  if (ge->exec2) CUDA_TRY(h, cudaGraphLaunch(ge->exec2, st));
  h->launches += ge->kernels;
  h->ctx_B = B;
  CUDA_TRY(h, cudaMemcpyAsync(x_inout, h->x, xb, out_kind, st));
  return 0;
}

MDTB200_API int mdtb200_sample_ancestral(MdtHandle* h, const float* sigmas, int n_steps, const float* goal, const float* state, int modality,
                                         int B, float* x_inout, const float* noise, float eta, void* stream) {
  if (!h) return MDTB200_EINVAL;
  TRY(check_sample_args(h, MDTB200_SAMPLER_EULER_ANCESTRAL, n_steps, B, sigmas, goal, state, x_inout));
  if (!noise) return fail(h, MDTB200_EINVAL, "sample_ancestral: null noise");
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->noise) {      // static (graph-visible) noise buffer: MAX_STEPS rows of max_batch * T * A draws, and the eta scalar
    h->noise_stride = (size_t)h->cfg.max_batch * h->T * h->A;
    int rc = 0;
    if ((rc = dev_alloc(h, &h->noise, (size_t)MAX_STEPS * h->noise_stride)) || (rc = dev_alloc(h, &h->anc_eta, 4))) return rc;
  }
  const size_t row = (size_t)B * h->T * h->A * 4;
  CUDA_TRY(h, cudaMemcpy2DAsync(h->noise, h->noise_stride * 4, noise, row, row, n_steps, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(h, cudaMemcpyAsync(h->anc_eta, &eta, 4, cudaMemcpyHostToDevice, st));      // 4 bytes from the stack: staged by the driver before return
  return sample_impl(h, MDTB200_SAMPLER_EULER_ANCESTRAL, sigmas, n_steps, goal, state, modality, B, x_inout, st, cudaMemcpyDeviceToDevice, cudaMemcpyDeviceToDevice);
}

MDTB200_API int mdtb200_sample(MdtHandle* h, int sampler, const float* sigmas, int n_steps, const float* goal, const float* state, int modality, int B,
                   float* x_inout, void* stream) {
  if (!h) return MDTB200_EINVAL;
  if (sampler == MDTB200_SAMPLER_EULER_ANCESTRAL) return fail(h, MDTB200_EINVAL, "the ancestral sampler needs the caller's noise: use mdtb200_sample_ancestral");
  return sample_impl(h, sampler, sigmas, n_steps, goal, state, modality, B, x_inout, (cudaStream_t)stream, cudaMemcpyDeviceToDevice, cudaMemcpyDeviceToDevice);
}

MDTB200_API int mdtb200_sample_host(MdtHandle* h, int sampler, const float* sigmas_host, int n_steps, const float* goal_host, const float* state_host, int modality,
                        int B, float* x_inout_host, void* stream) {
  if (!h) return MDTB200_EINVAL;
  if (sampler == MDTB200_SAMPLER_EULER_ANCESTRAL) return fail(h, MDTB200_EINVAL, "the ancestral sampler needs the caller's noise: use mdtb200_sample_ancestral");
  cudaStream_t st = (cudaStream_t)stream;
  TRY(sample_impl(h, sampler, sigmas_host, n_steps, goal_host, state_host, modality, B, x_inout_host, st, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return 0;
}

MDTB200_API int64_t mdtb200_launch_count(const MdtHandle* h) { return h ? h->launches : 0; }

MDTB200_API int64_t mdtb200_debug_copy(MdtHandle* h, const char* name, float* dst_dev, int64_t capacity, void* stream) {
  if (!h || !name || !dst_dev) return MDTB200_EINVAL;
  const size_t Bm = h->cfg.max_batch, M = Bm * h->T, d = h->d;
  struct Ent { const char* n; const float* p; size_t len; };
  const Ent tab[] = {
      {"ctx", h->ctx, M * d}, {"kv", h->kv, Bm * h->Tc * h->Ld * 2 * d}, {"mod", h->mod, (size_t)h->mod_rows * h->Ld * 6 * d},
      {"xh", h->xh, M * d}, {"xe", h->xe, M * d}, {"a", h->a, M * d}, {"qkv", h->qkv, M * 3 * d}, {"y", h->y, M * d},
      {"h", h->hbuf, M * 4 * d}, {"q", h->q, M * d}, {"pe", h->pe, (size_t)h->mod_rows * d}, {"cs", h->cs, (size_t)h->mod_rows * d},
      {"x", h->x, M * h->A},
      {"fused_trace", reinterpret_cast<const float*>(h->fd_trace), h->fd_trace ? (size_t)fd::TRACE_PHASES * 160 * 8 * 2 : 0},
  };
  for (const Ent& e : tab) {
    if (strcmp(e.n, name) == 0) {
      size_t n = e.len < (size_t)capacity ? e.len : (size_t)capacity;
      if (cudaMemcpyAsync(dst_dev, e.p, n * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream) != cudaSuccess) return fail(h, MDTB200_ECUDA, "debug_copy failed");
      return (int64_t)n;
    }
  }
  return fail(h, MDTB200_EINVAL, "debug_copy: unknown buffer '%s'", name);
}

// Kernel timeline (tools/ktrace.py): capacity > 0 arms the trace (allocates capacity records), capacity == 0 copies up to
// `max_records` {time, info} pairs to dst_host, disarms and frees.  Returns the number of records written (or a negative code).
MDTB200_API int64_t mdtb200_debug_ktrace(MdtHandle* h, int64_t capacity, unsigned long long* dst_host, int64_t max_records) {
  if (!h) return MDTB200_EINVAL;
  static unsigned long long* buf = nullptr;
  static unsigned int cap = 0;
  cudaDeviceSynchronize();
  if (capacity > 0) {
    if (buf) cudaFree(buf);
    if (cudaMalloc(&buf, (size_t)capacity * 16) != cudaSuccess) return fail(h, MDTB200_ENOMEM, "ktrace: allocation failed");
    cap = (unsigned int)capacity;
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_ktrace_n, &zero, sizeof(zero)); cudaMemcpyToSymbol(g_ktrace_cap, &cap, sizeof(cap)); cudaMemcpyToSymbol(g_ktrace, &buf, sizeof(buf));
    return 0;
  }
  unsigned int n = 0;
  unsigned long long* nullp = nullptr;
  cudaMemcpyFromSymbol(&n, g_ktrace_n, sizeof(n));
  cudaMemcpyToSymbol(g_ktrace, &nullp, sizeof(nullp));
  if (n > cap) n = cap;
  if ((int64_t)n > max_records) n = (unsigned int)max_records;
  if (buf && dst_host && n) cudaMemcpy(dst_host, buf, (size_t)n * 16, cudaMemcpyDeviceToHost);
  if (buf) { cudaFree(buf); buf = nullptr; }
  return (int64_t)n;
}

// Standalone tensor-core GEMM on fp32 inputs (tests only): splits A (M,K) and W (N,K) into bf16 hi|lo, runs the
// tcgen05 kernel with the handle's precision and writes fp32 out (M,N).  Synchronous; allocates scratch.
MDTB200_API int mdtb200_debug_gemm(MdtHandle* h, const float* A, const float* W, const float* bias, const float* R, const float* gate,
                                   int M, int N, int K, int epi, int rows_per_group, float* out, void* stream) {
  if (!h || !A || !W || !out) return MDTB200_EINVAL;
  if (h->cfg.precision == MDTB200_PREC_FP32) return fail(h, MDTB200_ESTATE, "debug_gemm needs a tensor-core precision handle");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16 *a16 = nullptr, *w16 = nullptr, *c16 = nullptr;
  const size_t Mp = (size_t)(M + 127) / 128 * 128;
  if (cudaMalloc(&a16, Mp * 2 * K * 2) != cudaSuccess || cudaMalloc(&w16, (size_t)N * 2 * K * 2) != cudaSuccess ||
      cudaMalloc(&c16, Mp * 2 * N * 2) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(a16); cudaFree(w16); cudaFree(c16);            // no scratch is leaked on a failed allocation
    return fail(h, MDTB200_ENOMEM, "debug_gemm: scratch allocation failed");
  }
  cudaMemsetAsync(a16, 0, Mp * 2 * K * 2, st);
  split_weights_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, a16, M, K);
  split_weights_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(W, w16, N, K);
  Gemm g;
  g.A16 = a16; g.lda16 = 2 * K; g.W16 = w16; g.bias = bias; g.C = out; g.ldc = N; g.C16 = c16; g.ldc16 = 2 * N; g.lo_off = N;
  g.R = R; g.ldr = N; g.gate = gate; g.gate_stride = N; g.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  g.M = M; g.N = N; g.K = K; g.epi = epi;
  unsigned long long* trace = nullptr;
  const bool want_trace = getenv("MDTB200_TRACE") != nullptr;
  const int n_cta = ((M + 127) / 128) * (N / 64);
  if (want_trace) { cudaMalloc(&trace, (size_t)n_cta * 8 * 8); cudaMemsetAsync(trace, 0, (size_t)n_cta * 8 * 8, st); }
  int rc = 0;
  for (int rep = 0; rep < (want_trace ? 3 : 1) && !rc; ++rep) {   // trace the 3rd (warm) launch
    tc::TcGemm t{};
    t.A16 = g.A16; t.lda16 = g.lda16; t.W16 = g.W16; t.bias = g.bias; t.C = g.C; t.ldc = g.ldc; t.C16 = g.C16; t.ldc16 = g.ldc16; t.lo_off = g.lo_off;
    t.R = g.R; t.ldr = g.ldr; t.gate = g.gate; t.gate_stride = g.gate_stride; t.rows_per_group = g.rows_per_group;
    t.M = M; t.N = N; t.K = K; t.epi = epi; t.passes = h->cfg.precision == MDTB200_PREC_BF16X3 ? 3 : 1; t.trace = trace;
    t.w_dynamic = 1;     // W was split by the kernel launched just before
    const char* em = tc::launch_tc_gemm(h->tma, t, st);
    if (em) rc = fail(h, MDTB200_ECUDA, "tcgen05 gemm: %s", em);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (want_trace && !rc && e == cudaSuccess) {
    std::vector<unsigned long long> tr((size_t)n_cta * 8);
    cudaMemcpy(tr.data(), trace, tr.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull, t6 = 0;
    for (int c = 0; c < n_cta; ++c) if (tr[c * 8]) { if (tr[c * 8] < t0) t0 = tr[c * 8]; if (tr[c * 8 + 7] > t6) t6 = tr[c * 8 + 7]; }
    fprintf(stderr, "[trace] M=%d N=%d K=%d epi=%d: grid span %.2f us; per-CTA ns since first CTA start: entry/setup/first-full/last-full/acc-ready/phaseA-done/sync-done/phaseB-done\n", M, N, K, epi, (t6 - t0) / 1e3);
    for (int c = 0; c < n_cta; c += (n_cta > 6 ? n_cta / 6 : 1)) {
      if (!tr[c * 8]) continue;
      fprintf(stderr, "[trace]   cta %4d:", c);
      for (int j = 0; j < 8; ++j) fprintf(stderr, " %7lld", (long long)(tr[c * 8 + j] - t0));
      fprintf(stderr, "\n");
    }
  }
  if (trace) cudaFree(trace);
  h->tma.cache.clear();
  cudaFree(a16); cudaFree(w16); cudaFree(c16);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, MDTB200_ECUDA, "debug_gemm: %s", cudaGetErrorString(e));
  return 0;
}

// Times `iters` back-to-back launches of the tensor-core GEMM kernel (pseudo-random operands, pre-split, L2-warm) with CUDA
// events on `stream`; used by bench.py for the per-kernel roofline entry.  epi as in mdtb200_debug_gemm.
MDTB200_API int mdtb200_debug_gemm_time(MdtHandle* h, int M, int N, int K, int epi, int iters, float* avg_us, void* stream) {
  if (!h || !avg_us || iters < 1) return MDTB200_EINVAL;
  if (h->cfg.precision == MDTB200_PREC_FP32) return fail(h, MDTB200_ESTATE, "debug_gemm_time needs a tensor-core precision handle");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Mp = (size_t)(M + 127) / 128 * 128;
  __nv_bfloat16 *a16 = nullptr, *w16 = nullptr, *c16 = nullptr; float *c = nullptr, *gate = nullptr;
  if (cudaMalloc(&a16, Mp * 2 * K * 2) != cudaSuccess || cudaMalloc(&w16, (size_t)N * 2 * K * 2) != cudaSuccess ||
      cudaMalloc(&c16, Mp * 2 * N * 2) != cudaSuccess || cudaMalloc(&c, Mp * N * 4) != cudaSuccess || cudaMalloc(&gate, Mp * N * 4) != cudaSuccess) {
    cudaGetLastError();
    cudaFree(a16); cudaFree(w16); cudaFree(c16); cudaFree(c); cudaFree(gate);
    return fail(h, MDTB200_ENOMEM, "debug_gemm_time: scratch allocation failed");
  }
  // pseudo-random, non-zero split-bf16 operands (hi ~ U(-1, 1), lo ~ 2^-9 of that), as the graph's GEMMs see them
  fill_operand_kernel<<<(unsigned)((Mp * K + 255) / 256), 256, 0, st>>>(a16, (int64_t)Mp, K, 0x1234u);
  fill_operand_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(w16, (int64_t)N, K, 0x9876u);
  cudaMemsetAsync(c, 0, Mp * N * 4, st); cudaMemsetAsync(gate, 0, Mp * N * 4, st);
  tc::TcGemm t{};
  t.A16 = a16; t.lda16 = 2 * K; t.W16 = w16; t.C = (epi == EPI_GELU) ? nullptr : c; t.ldc = N;
  t.C16 = (epi == EPI_GELU) ? c16 : nullptr; t.ldc16 = 2 * N; t.lo_off = N;
  t.R = c; t.ldr = N; t.gate = gate; t.gate_stride = 0; t.rows_per_group = 10;
  t.M = M; t.N = N; t.K = K; t.epi = epi; t.passes = h->cfg.precision == MDTB200_PREC_BF16X3 ? 3 : 1; t.trace = nullptr;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* em = nullptr;
  for (int i = 0; i < 5 && !em; ++i) em = tc::launch_tc_gemm(h->tma, t, st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters && !em; ++i) em = tc::launch_tc_gemm(h->tma, t, st);
  cudaEventRecord(e1, st);
  cudaError_t e = cudaStreamSynchronize(st);
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  *avg_us = ms * 1000.f / iters;
  h->launches += iters + 5;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  h->tma.cache.clear();
  cudaFree(a16); cudaFree(w16); cudaFree(c16); cudaFree(c); cudaFree(gate);
  if (em) return fail(h, MDTB200_ECUDA, "debug_gemm_time: %s", em);
  if (e != cudaSuccess) return fail(h, MDTB200_ECUDA, "debug_gemm_time: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"

#include "ops_train2.cuh"
#include "perceiver_host.cuh"
