// Host orchestration + C ABI of the PerceiverResampler (kernels: perceiver.cuh).  Included by engine.cu (single translation unit).
#pragma once
#include "perceiver.cuh"

struct MdtPerceiver {
  MdtPerceiverConfig cfg;
  int device = 0;
  int d = 0, H = 0, inner = 0, Q = 0, depth = 0, ff = 0;
  std::string err;
  std::map<std::string, Bound> bound;
  bool committed = false;
  std::vector<void*> allocs;
  tc::TmaEncoder tma;
  int64_t launches = 0;

  struct Layer {
    const float *nl_w, *nl_b, *ff_w, *ff_b, *kb, *vb, *gm, *wq, *wk;
    const __nv_bfloat16 *wv16, *wvx16, *wout16, *w1_16, *w2_16;
  };
  std::vector<Layer> layers;
  const float *latents = nullptr, *tpe = nullptr, *norm_w = nullptr, *norm_b = nullptr;
  float* arena = nullptr; size_t arena_floats = 0, arena_used = 0;
  __nv_bfloat16* arena16 = nullptr; size_t arena16_elems = 0, arena16_used = 0;

  // workspace (max_batch samples, max_features feature tokens)
  int Mp = 0, Fp = 0;
  float *xhat = nullptr, *L = nullptr, *lnf = nullptr, *qkv = nullptr, *qt = nullptr, *cq = nullptr, *scores = nullptr, *wsum = nullptr, *olat = nullptr, *of = nullptr;
  __nv_bfloat16 *a16 = nullptr, *z16 = nullptr, *o16 = nullptr, *h16 = nullptr;
};

namespace {

int pfail(MdtPerceiver* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return code;
}
#define PCUDA(h, expr)                                                                                \
  do {                                                                                                \
    cudaError_t e_ = (expr);                                                                          \
    if (e_ != cudaSuccess) return pfail(h, MDTB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

template <typename T>
int palloc(MdtPerceiver* h, T** p, size_t n) {
  void* q = nullptr;
  if (cudaMalloc(&q, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return pfail(h, MDTB200_ENOMEM, "cudaMalloc(%zu bytes) failed", n * sizeof(T)); }
  h->allocs.push_back(q);
  *p = reinterpret_cast<T*>(q);
  return 0;
}

int pcheck(MdtPerceiver* h, const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return pfail(h, MDTB200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e)); }
  h->launches++;
  return 0;
}

int pgemm(MdtPerceiver* h, tc::TcGemm t, cudaStream_t st) {
  t.passes = 3; t.rows_per_group = 1; t.trace = nullptr;
  const char* e = tc::launch_tc_gemm(h->tma, t, st);
  if (e) return pfail(h, MDTB200_ECUDA, "perceiver gemm (M=%d N=%d K=%d): %s", t.M, t.N, t.K, e);
  return pcheck(h, "tc_gemm_kernel");
}

int pln(MdtPerceiver* h, const float* x, float* out, __nv_bfloat16* out16, const float* w, const float* b, int M, cudaStream_t st) {
  LnArgs a{};
  a.x = x; a.out = out; a.out16 = out16; a.ld16 = 2 * h->d; a.lo_off = h->d; a.w = w; a.b = b; a.rows_per_group = 1; a.M = M; a.d = h->d;
  const int blocks = (M * 32 + 255) / 256;
  launch_pdl(ln_mod_kernel<3>, dim3(blocks), dim3(256), 0, st, a);
  return pcheck(h, "ln_mod_kernel");
}

struct PPacker {
  MdtPerceiver* h; cudaStream_t st; bool ok = true;
  const float* src(const std::string& name, int64_t numel) {
    auto it = h->bound.find(name);
    if (it == h->bound.end()) { ok = false; pfail(h, MDTB200_ESTATE, "weight '%s' was not bound", name.c_str()); return nullptr; }
    if (it->second.numel != numel) { ok = false; pfail(h, MDTB200_EINVAL, "weight '%s': expected %lld elements, got %lld", name.c_str(), (long long)numel, (long long)it->second.numel); return nullptr; }
    return it->second.ptr;
  }
  float* take(size_t n) {
    if (h->arena_used + n > h->arena_floats) { ok = false; pfail(h, MDTB200_ENOMEM, "perceiver weight arena overflow"); return nullptr; }
    float* r = h->arena + h->arena_used;
    h->arena_used += (n + 31) / 32 * 32;
    return r;
  }
  __nv_bfloat16* take16(size_t n) {
    if (h->arena16_used + n > h->arena16_elems) { ok = false; pfail(h, MDTB200_ENOMEM, "perceiver bf16 arena overflow"); return nullptr; }
    __nv_bfloat16* r = h->arena16 + h->arena16_used;
    h->arena16_used += (n + 63) / 64 * 64;
    return r;
  }
  const float* copy(const std::string& name, int64_t numel) {
    const float* s = src(name, numel);
    float* dst = ok ? take((size_t)numel) : nullptr;
    if (s && dst) cudaMemcpyAsync(dst, s, (size_t)numel * 4, cudaMemcpyDeviceToDevice, st);
    return dst;
  }
  const __nv_bfloat16* split(const float* w, int64_t rows, int cols) {
    if (!w || !ok) return nullptr;
    __nv_bfloat16* dst = take16((size_t)rows * cols * 2);
    if (dst) split_weights_kernel<<<(unsigned)((rows * cols + 255) / 256), 256, 0, st>>>(w, dst, rows, cols);
    return dst;
  }
};

}  // namespace

extern "C" {

MDTB200_API const char* mdtb200_perceiver_last_error(const MdtPerceiver* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

MDTB200_API void mdtb200_perceiver_destroy(MdtPerceiver* h) {
  if (!h) return;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
}

MDTB200_API int mdtb200_perceiver_create(const MdtPerceiverConfig* cfg, MdtPerceiver** out) {
  if (!cfg || !out) return pfail(nullptr, MDTB200_EINVAL, "null argument");
  *out = nullptr;
  if (cfg->abi_version != MDTB200_ABI_VERSION) return pfail(nullptr, MDTB200_EINVAL, "abi_version %d != %d", cfg->abi_version, MDTB200_ABI_VERSION);
  const int d = cfg->dim;
  if (d != 384) return pfail(nullptr, MDTB200_EUNSUPPORTED, "perceiver dim %d unsupported (384: the shipped MDT-V configuration)", d);
  if (cfg->dim_head != 64) return pfail(nullptr, MDTB200_EUNSUPPORTED, "perceiver dim_head %d unsupported (64)", cfg->dim_head);
  if (cfg->heads < 1 || cfg->num_latents < 1 || cfg->heads * cfg->num_latents > 64) return pfail(nullptr, MDTB200_EUNSUPPORTED, "heads * num_latents must be <= 64");
  if (cfg->depth < 1 || cfg->ff_mult != 4 || cfg->num_time_embeds < 1 || cfg->max_batch < 1 || cfg->max_features < 1)
    return pfail(nullptr, MDTB200_EINVAL, "bad perceiver configuration (depth >= 1, ff_mult == 4, num_time_embeds / max_batch / max_features >= 1)");
  MdtPerceiver* h = new (std::nothrow) MdtPerceiver();
  if (!h) return pfail(nullptr, MDTB200_ENOMEM, "out of host memory");
  h->cfg = *cfg;
  h->d = d; h->H = cfg->heads; h->inner = cfg->heads * 64; h->Q = cfg->num_latents; h->depth = cfg->depth; h->ff = 4 * d;
  auto bail = [&](int code) { g_create_error = h->err; mdtb200_perceiver_destroy(h); return code; };
  if (cudaGetDevice(&h->device) != cudaSuccess) { pfail(h, MDTB200_ECUDA, "no CUDA device"); return bail(MDTB200_ECUDA); }
  cudaDeviceProp prop{};
  cudaGetDeviceProperties(&prop, h->device);
  if (prop.major != 10) { pfail(h, MDTB200_EUNSUPPORTED, "the perceiver needs an sm_100 device (found sm_%d%d)", prop.major, prop.minor); return bail(MDTB200_EUNSUPPORTED); }
  if (const char* e = h->tma.init()) { pfail(h, MDTB200_ECUDA, "TMA descriptor encoder unavailable: %s", e); return bail(MDTB200_ECUDA); }
  if (const char* e = tc::configure_kernels()) { pfail(h, MDTB200_ECUDA, "tcgen05 kernel configuration failed: %s", e); return bail(MDTB200_ECUDA); }
  const size_t D = d, I = h->inner, FF = h->ff, HQ = (size_t)h->H * h->Q;
  const size_t ss = pr::score_smem_bytes((int)HQ, d), zs = pr::z_smem_bytes((int)HQ, cfg->max_features, h->Q);
  if (ss > 200 * 1024 || zs > 200 * 1024) { pfail(h, MDTB200_EUNSUPPORTED, "max_features %d too large for the attention kernels", cfg->max_features); return bail(MDTB200_EUNSUPPORTED); }
  if (pr::attn_configure((int)HQ, ss, zs) != cudaSuccess) {
    pfail(h, MDTB200_ECUDA, "perceiver attention kernels need %zu / %zu bytes of shared memory", ss, zs); return bail(MDTB200_ECUDA);
  }
  int rc = 0;
  h->arena_floats = (size_t)h->Q * D + (size_t)cfg->num_time_embeds * D + 2 * D + h->depth * (8 * D + 2 * I + 2 * I * D + 1024) + (1 << 14);
  h->arena16_elems = (size_t)h->depth * 2 * (I * D + I * D + D * I + 2 * FF * D) + (1 << 14);
  if ((rc = palloc(h, &h->arena, h->arena_floats)) || (rc = palloc(h, &h->arena16, h->arena16_elems))) return bail(rc);
  const size_t B = cfg->max_batch, F = cfg->max_features;
  h->Mp = (int)((B * h->Q + 127) / 128 * 128);
  h->Fp = (int)((F + 31) / 32 * 32);
  const size_t Mp = h->Mp;
  if ((rc = palloc(h, &h->xhat, B * F * D)) || (rc = palloc(h, &h->L, Mp * D)) || (rc = palloc(h, &h->lnf, Mp * D)) || (rc = palloc(h, &h->qkv, Mp * 3 * I)) ||
      (rc = palloc(h, &h->qt, h->H * Mp * D)) || (rc = palloc(h, &h->cq, Mp * h->H)) || (rc = palloc(h, &h->scores, B * HQ * h->Fp)) ||
      (rc = palloc(h, &h->wsum, Mp * h->H)) || (rc = palloc(h, &h->olat, Mp * I)) || (rc = palloc(h, &h->of, h->H * Mp * 64)) ||
      (rc = palloc(h, &h->a16, Mp * 2 * D)) || (rc = palloc(h, &h->z16, h->H * Mp * 2 * D)) ||
      (rc = palloc(h, &h->o16, Mp * 2 * I)) || (rc = palloc(h, &h->h16, Mp * 2 * FF)))
    return bail(rc);
  cudaMemset(h->a16, 0, Mp * 2 * D * 2); cudaMemset(h->z16, 0, h->H * Mp * 2 * D * 2); cudaMemset(h->o16, 0, Mp * 2 * I * 2);
  cudaMemset(h->h16, 0, Mp * 2 * FF * 2); cudaMemset(h->L, 0, Mp * D * 4);
  *out = h;
  return 0;
}

MDTB200_API int mdtb200_perceiver_bind_weight(MdtPerceiver* h, const char* name, const float* dev_ptr, int64_t numel) {
  if (!h) return MDTB200_EINVAL;
  if (!name || !dev_ptr || numel <= 0) return pfail(h, MDTB200_EINVAL, "bind_weight: bad argument");
  h->bound[name] = Bound{dev_ptr, numel};
  return 0;
}

// Packs the bound state dict (reference key names, perceiver_resampler.py:86-117): fused [Wq * scale; Wk; Wv], the norm_media affine
// folded into per-head key / value operands, Wk b and Wv b, split-bf16 copies of every GEMM weight.
MDTB200_API int mdtb200_perceiver_commit_weights(MdtPerceiver* h, void* stream) {
  if (!h) return MDTB200_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  h->committed = false; h->arena_used = 0; h->arena16_used = 0; h->tma.cache.clear();
  const int64_t d = h->d, I = h->inner, FF = h->ff;
  PPacker pk{h, st};
  h->latents = pk.copy("latents", (int64_t)h->Q * d);
  h->tpe = pk.copy("time_pos_emb", (int64_t)h->cfg.num_time_embeds * d);
  h->layers.assign(h->depth, MdtPerceiver::Layer{});
  for (int l = 0; l < h->depth && pk.ok; ++l) {
    MdtPerceiver::Layer& L = h->layers[l];
    const std::string a = "layers." + std::to_string(l) + ".0.", f = "layers." + std::to_string(l) + ".1.";
    L.gm = pk.copy(a + "norm_media.weight", d);
    const float* bm = pk.copy(a + "norm_media.bias", d);
    L.nl_w = pk.copy(a + "norm_latents.weight", d); L.nl_b = pk.copy(a + "norm_latents.bias", d);
    L.ff_w = pk.copy(f + "0.weight", d); L.ff_b = pk.copy(f + "0.bias", d);
    L.wq = pk.copy(a + "to_q.weight", I * d); L.wk = pk.copy(a + "to_k.weight", I * d);      // fp32: exact score path
    const float *wv = pk.src(a + "to_v.weight", I * d), *wo = pk.src(a + "to_out.weight", d * I);
    const float *w1 = pk.src(f + "1.weight", FF * d), *w2 = pk.src(f + "3.weight", d * FF);
    float *kb = pk.take((size_t)I), *vb = pk.take((size_t)I);
    if (!pk.ok || !kb || !vb) break;
    L.wv16 = pk.split(wv, I, (int)d);
    __nv_bfloat16* wvx = pk.take16((size_t)I * 2 * d);
    if (!pk.ok) break;
    pr::perceiver_fold_kernel<<<(unsigned)((h->H * 64 * d + 255) / 256), 256, 0, st>>>(wv, L.gm, wvx, (int)d, h->H);
    L.wvx16 = wvx;
    {   // kb = Wk b_media, vb = Wv b_media (one row per output: the skinny kernel with M = 1)
      SkinnyArgs g1{bm, L.wk, nullptr, kb, 1, (int)I, (int)d, EPI_NONE}, g2{bm, wv, nullptr, vb, 1, (int)I, (int)d, EPI_NONE};
      skinny_gemm_kernel<<<(unsigned)((I + 7) / 8), 256, (size_t)d * 4, st>>>(g1);
      skinny_gemm_kernel<<<(unsigned)((I + 7) / 8), 256, (size_t)d * 4, st>>>(g2);
    }
    L.kb = kb; L.vb = vb;
    L.wout16 = pk.split(wo, d, (int)I); L.w1_16 = pk.split(w1, FF, (int)d); L.w2_16 = pk.split(w2, d, (int)FF);
  }
  h->norm_w = pk.copy("norm.weight", d); h->norm_b = pk.copy("norm.bias", d);
  h->bound.clear();
  if (!pk.ok) return MDTB200_ESTATE;
  PCUDA(h, cudaGetLastError());
  PCUDA(h, cudaStreamSynchronize(st));
  h->committed = true;
  return 0;
}

// PerceiverResampler.forward (perceiver_resampler.py:126-163): x_f (B, T, n, d) device, mask (B, T) floats or NULL -> out (B, Q, d)
MDTB200_API int mdtb200_perceiver_forward(MdtPerceiver* h, const float* x_f, const float* mask, int B, int T, int n, float* out, void* stream) {
  if (!h) return MDTB200_EINVAL;
  if (!h->committed) return pfail(h, MDTB200_ESTATE, "weights not committed (call mdtb200_perceiver_commit_weights)");
  if (!x_f || !out) return pfail(h, MDTB200_EINVAL, "perceiver_forward: null argument");
  const int F = T * n;
  if (B < 1 || B > h->cfg.max_batch || T < 1 || T > h->cfg.num_time_embeds || F > h->cfg.max_features)
    return pfail(h, MDTB200_EINVAL, "perceiver_forward: B=%d T=%d n=%d outside the configured limits (max_batch %d, num_time_embeds %d, max_features %d)",
                 B, T, n, h->cfg.max_batch, h->cfg.num_time_embeds, h->cfg.max_features);
  int dev = -1;
  cudaGetDevice(&dev);
  if (dev != h->device) return pfail(h, MDTB200_ESTATE, "handle belongs to device %d but device %d is current", h->device, dev);
  cudaStream_t st = (cudaStream_t)stream;
  const int d = h->d, H = h->H, I = h->inner, Q = h->Q, FF = h->ff, Mq = B * Q, Mp = (Mq + 127) / 128 * 128, Fp = h->Fp, HQ = H * Q;
  {
    pr::PrepArgs a{x_f, h->tpe, mask, h->xhat, B, T, n, d};
    const int blocks = (B * F * 32 + 255) / 256;
    launch_pdl(pr::perceiver_prep_kernel<3>, dim3(blocks), dim3(256), 0, st, a);
    TRY(pcheck(h, "perceiver_prep_kernel"));
    launch_pdl(pr::perceiver_init_kernel, dim3((Mq * d + 255) / 256), dim3(256), 0, st, h->latents, h->L, B, Q, d);
    TRY(pcheck(h, "perceiver_init_kernel"));
  }
  for (int l = 0; l < h->depth; ++l) {
    const MdtPerceiver::Layer& L = h->layers[l];
    TRY(pln(h, h->L, h->lnf, h->a16, L.nl_w, L.nl_b, Mq, st));
    {   // exact-fp32 score path: q (scaled), latent keys, feature-space queries, score constants
      pr::QPathArgs a{h->lnf, L.wq, L.wk, L.gm, L.kb, h->qkv, 3 * I, I, h->qt, h->cq, Mq, Mp, H, d, 1.0f / sqrtf(64.0f)};
      static const bool simt = getenv("MDTB200_PERC_QPATH") && !strcmp(getenv("MDTB200_PERC_QPATH"), "simt");
      if (simt) {
        launch_pdl(pr::perceiver_qpath_kernel<384>, dim3((Mq + pr::QP_ROWS - 1) / pr::QP_ROWS, H), dim3(256), 0, st, a);
      } else {      // 3xTF32 tensor-core kernels: [q | k] projection (+ cq), then the feature-space queries
        launch_pdl(pr::perceiver_qk_mma_kernel, dim3((Mq + 63) / 64, 2 * H), dim3(128), 0, st, a);
        launch_pdl(pr::perceiver_qt_mma_kernel, dim3((Mq + 63) / 64, d / 128, H), dim3(128), 0, st, a);
        h->launches += 1;
      }
      TRY(pcheck(h, "perceiver q-path kernels"));
    }
    {   // v_lat = LN(latents) . Wv^T  (value path: tensor cores)
      tc::TcGemm t{};
      t.A16 = h->a16; t.lda16 = 2 * d; t.W16 = L.wv16; t.C = h->qkv + 2 * I; t.ldc = 3 * I; t.M = Mq; t.N = I; t.K = d; t.epi = EPI_NONE;
      TRY(pgemm(h, t, st));
    }
    {
      pr::ScoreArgs a{h->qt, h->cq, h->xhat, h->scores, B, F, Fp, Q, H, Mp, d};
      pr::ZArgs z{h->scores, h->qkv, 3 * I, I, h->xhat, h->z16, h->wsum, h->olat, B, F, Fp, Q, H, Mp, d};
      pr::attn_launch(a, z, st);
      TRY(pcheck(h, "perceiver attention kernels"));
    }
    {   // Wv_h (g (.) z_h) : head-grouped GEMM, N = 64
      tc::TcGemm t{};
      t.A16 = h->z16; t.lda16 = 2 * d; t.W16 = L.wvx16; t.w_rows = H * 64; t.wg_rows = Mp; t.wg_stride = 64;
      t.C = h->of; t.ldc = 64; t.M = H * Mp; t.N = 64; t.K = d; t.epi = EPI_NONE;
      TRY(pgemm(h, t, st));
      pr::CombineArgs c{h->of, h->wsum, L.vb, h->olat, h->o16, Mq, Mp, H, I};
      launch_pdl(pr::perceiver_combine_kernel, dim3((Mq * I + 255) / 256), dim3(256), 0, st, c);
      TRY(pcheck(h, "perceiver_combine_kernel"));
    }
    {   // latents += to_out(o)
      tc::TcGemm t{};
      t.A16 = h->o16; t.lda16 = 2 * I; t.W16 = L.wout16; t.C = h->L; t.ldc = d; t.R = h->L; t.ldr = d; t.M = Mq; t.N = d; t.K = I; t.epi = EPI_RES;
      TRY(pgemm(h, t, st));
    }
    TRY(pln(h, h->L, nullptr, h->a16, L.ff_w, L.ff_b, Mq, st));
    {   // latents += W2 gelu(W1 LN(latents))
      tc::TcGemm t{};
      t.A16 = h->a16; t.lda16 = 2 * d; t.W16 = L.w1_16; t.C16 = h->h16; t.ldc16 = 2 * FF; t.lo_off = FF; t.M = Mq; t.N = FF; t.K = d; t.epi = EPI_GELU;
      TRY(pgemm(h, t, st));
      tc::TcGemm u{};
      u.A16 = h->h16; u.lda16 = 2 * FF; u.W16 = L.w2_16; u.C = h->L; u.ldc = d; u.R = h->L; u.ldr = d; u.M = Mq; u.N = d; u.K = FF; u.epi = EPI_RES;
      TRY(pgemm(h, u, st));
    }
  }
  return pln(h, h->L, out, nullptr, h->norm_w, h->norm_b, Mq, st);
}

MDTB200_API int64_t mdtb200_perceiver_launch_count(const MdtPerceiver* h) { return h ? h->launches : 0; }

}  // extern "C"
