// Persistent fused decoder ("phase machine") for sm_100a: the whole N-step sampling loop of the score network in ONE launch.
//
// Why: at B=256 the decoder is a chain of ~46 dependent kernels per score evaluation, each 30-60 CTAs and ~8 us of mostly
// latency (profiles/r01_v7_summary.md).  Samples never interact, so the batch is cut into ROW GROUPS of SPG = floor(128 / T)
// samples (120 token rows = one 128-row UMMA tile) and each row group is owned by a GROUP of C = d / 64 CTAs for the whole
// call.  The CTAs of a group execute the same PROGRAM (array of Phase descriptors built by the host, engine.cu):
//
//   GEMM phase   every CTA computes a column slice of D = A . W^T for the group's 128 rows: A (bf16 hi|lo operand written by an
//                earlier phase) and its W slice stream through a 2-stage TMA ring that persists across phases (the first W
//                tile of the next GEMM is requested while the current epilogue / row phase still runs), tcgen05.mma
//                (M=128, N = 64/192/256, bf16x3) accumulates in one of two TMEM buffers, 16 worker warps run the epilogue
//                (bias, GELU, gated residual, split-bf16 store, split-K partial store) through a coalescing staging tile.
//   ATTN phase   softmax(q k^T / sqrt(hd) + causal top-left mask) v for the CTA's share of the group's samples (SIMT, smem).
//   ROW phase    warp per token row: [sum of split-K partials + bias -> gated residual] -> [LN_dec -> action_pred -> EDM
//                preconditioner -> sampler update -> next evaluation's action embedding] -> [LayerNorm (+AdaLN modulate) ->
//                split-bf16 operand of the next GEMM].
//
// Phases exchange activations through L2 (global buffers).  Synchronisation is per group: every CTA publishes "phases
// completed" with st.release.gpu, consumers poll the C counters of their group with ld.acquire.gpu (bounded spin -> trap), so
// there is no grid-wide barrier and no kernel boundary inside a sampling call.  Reference math: transformer_blocks.py:292-309
// (ConditionedBlock), mdtv_transformer.py:224-236 (forward_dec_only), gc_sampling.py (samplers), score_wrappers.py:31-43.
#pragma once
#include "gemm_tcgen05.cuh"

namespace mdt { namespace fd {

using namespace tc;

constexpr int FD_WARPS = 20;                 // warps 0..15: workers (4 warpgroups); warp 16: TMA producer, warp 17: MMA issuer, 18-19 idle
constexpr int FD_THREADS = FD_WARPS * 32;    // 640: setmaxnreg moves registers from the last warpgroup (32/thread) to the workers (112)
constexpr int FD_WORKERS = 512;
constexpr int A_TILE = BM * BK * 2;          // 16 KB: 128 rows x 64 bf16, SWIZZLE_128B
constexpr int W_TILE_MAX = 256 * BK * 2;     // 32 KB: up to 256 weight rows
constexpr int STAGE = 2 * A_TILE + 2 * W_TILE_MAX;   // A hi | A lo | W hi | W lo = 96 KB
constexpr int NSTAGE = 2;
constexpr int STG_BYTES = BM * 64 * 4;       // epilogue staging tile: 128 rows x 64 floats, 16-byte slots XOR-swizzled by row
constexpr int EPI_VEC = 256;                 // bias / gate slices of the current GEMM phase (floats each)
constexpr int SCR_OFF = NSTAGE * STAGE;
constexpr int SCR_BYTES = STG_BYTES + 2 * EPI_VEC * 4;   // 34,816
constexpr int BAR_OFF = SCR_OFF + SCR_BYTES;
constexpr int PH_OFF = BAR_OFF + 256;                // shared-memory copy of the workers' current Phase (512 bytes)
constexpr int SMEM_TOTAL = PH_OFF + 512;             // 232,192 <= 232,448
constexpr int ROW_SCR_OFF = STAGE;                   // row / attention phases use ring stage 1 + the staging tile
constexpr int ROW_SCR_BYTES = STAGE + SCR_BYTES;     // 133,120
constexpr int ACC_COLS = 256;                        // TMEM columns per accumulator buffer (2 buffers = 512)

enum : int { PH_END = 0, PH_GEMM = 1, PH_ATTN = 2, PH_ROW = 3 };
enum : int { FE_STORE = 0, FE_RESID = 1, FE_GELU16 = 2, FE_PARTIAL = 3 };

struct alignas(16) Phase {
  int type;
  int dep;                       // phases every CTA of the group must have completed before this one reads its inputs
  // ---------------------------------------------------------------- GEMM, worker side (epilogue); operand side: GemmDesc
  // per-CTA offsets: X = X_base + (cta % cta_mod) * X_s1 + (cta / cta_mod) * X_s2
  int bn, epi, acc_buf, cta_mod;
  int w_row_base, w_row_s1, w_row_s2;      // bias index = W row of this CTA's slice
  int o_col_base, o_col_s1, o_col_s2, ldo, lo_off16;
  long long out_cta_stride;      // floats added to `out` per (cta % cta_mod) (split-K partial buffers)
  const float* bias;             // indexed by W row (may be null)
  const float* gate; int gate_stride;   // FE_RESID: gate[(row / T) * gate_stride + col] (null: plain residual)
  float* out; __nv_bfloat16* out16;
  // ---------------------------------------------------------------- ATTN (out16 / ldo / lo_off16 shared with GEMM)
  const float *q, *k, *v; int ldq, ldkv, Tq, Tk, causal; float att_scale;
  // ---------------------------------------------------------------- ROW
  int n_part; const float* part; long long part_stride; const float* pbias;   // split-K partial sum (+bias)
  const float* rgate; int rgate_stride;                                        // gated residual into xh
  float* xh;                                                                   // residual stream (M, d)
  int head_mode;                 // -1: no head; else HeadMode of kernels_simt.cuh
  const float *dln_w, *dln_b, *ap_w, *ap_b;
  const float* x_in; float* x_state; float* x_aux; float* dbuf; float* hout;
  const float* sigmas; int step, n_steps; const float* hsigma; int hsigma_stride;
  int embed;                     // 1: xh = action_emb(x * c_in) of the NEXT evaluation; emb_x null: x comes from the head
  const float* emb_x; const float* ae_w; const float* ae_b; const float* emb_sigma; int emb_sigma_stride; int precondition;
  int ln;                        // 1: LayerNorm (+modulate) -> out16
  const float *ln_w, *ln_b, *shift, *scale; int mod_stride;
};

// Operand side of a GEMM phase, read by the TMA producer and the MMA issuer (kept in registers; 80 bytes)
struct GemmDesc {
  int p, dep, a_map, w_map, nkb, bn, acc_buf, cta_mod;
  int a_col_base, a_col_s1, a_lo_off;                  // A column (k) offset = base + (cta % cta_mod) * s1; lo half at +a_lo_off
  int w_row_base, w_row_s1, w_row_s2, w_col_base, w_col_s1, w_lo_off;
  int pad0, pad1, pad2;
};
static_assert(sizeof(GemmDesc) == 80, "GemmDesc layout");
static_assert(sizeof(Phase) <= 512 && sizeof(Phase) % 16 == 0, "Phase must fit the 512-byte shared-memory slot (one uint4 per lane)");

struct FusedParams {
  const Phase* prog; const GemmDesc* gemms; int n_gemms; const CUtensorMap* maps; int* progress;   // progress[row group][32] ints, zeroed before the launch
  int B, T, A, d, H, hd, C, SPG, n_rowgroups, passes;
  float sigma_data;
  unsigned long long* trace;     // optional (MDTB200_FUSED_TRACE): globaltimer stamps [phase < TRACE_PHASES][cta][8]
};
constexpr int TRACE_PHASES = 128;
#define FD_STAMP(p, slot) do { if (P.trace && (p) < TRACE_PHASES) P.trace[((size_t)(p) * gridDim.x + blockIdx.x) * 8 + (slot)] = gtimer(); } while (0)

// ------------------------------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_workers() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// All lanes of the calling warp wait until every CTA of the group has completed `dep` phases (bounded: a protocol bug traps).
// Polling uses relaxed loads (served by L2, no L1 invalidation per iteration); ONE acquire fence follows the successful poll.
__device__ __forceinline__ void group_wait(const int* prow, int C, int dep, int lane) {
  if (dep <= 0) return;
  const long long t0 = clock64();
  for (;;) {
    int ok = 1;
    if (lane < C) ok = ld_relaxed_gpu(prow + lane) >= dep;
    if (__all_sync(0xffffffffu, ok)) break;
    if (clock64() - t0 > 8000000000LL) __trap();
  }
  fence_acq_rel_gpu();
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------ sampler update (head)
// Same fp32 op order as head_kernel (kernels_simt.cuh); returns the value the NEXT evaluation sees for element e.
__device__ __forceinline__ float head_update(const Phase& ph, int mode, size_t e, int row, int T, float raw, float sigma_data) {
  if (mode == HEAD_RAW) { ph.hout[e] = raw; return 0.f; }
  float sig, sig_next = 0.f;
  if (mode <= HEAD_DENOISE) {
    sig = ph.hsigma ? ph.hsigma[(size_t)(row / T) * ph.hsigma_stride] : 1.f;
  } else {
    sig = ph.sigmas[mode == HEAD_HEUN2 ? ph.step + 1 : ph.step];
    sig_next = ph.sigmas[ph.step + 1];
  }
  float c_skip, c_out, c_in;
  edm_scalings(sig, sigma_data, c_skip, c_out, c_in);
  const float xin = __ldcg(ph.x_in + e);
  const float D = raw * c_out + xin * c_skip;
  float nx = 0.f;
  switch (mode) {
    case HEAD_DENOISE: ph.hout[e] = D; break;
    case HEAD_DDIM: {
      float t = -logf(sig), tn = -logf(sig_next), h = tn - t;
      nx = (expf(-tn) / expf(-t)) * xin - expm1f(-h) * D;
      ph.x_state[e] = nx;
    } break;
    case HEAD_EULER: {
      float dd = (xin - D) / sig;
      nx = xin + dd * (sig_next - sig);
      ph.x_state[e] = nx;
    } break;
    case HEAD_HEUN1: {
      float dd = (xin - D) / sig, dt = sig_next - sig;
      nx = xin + dd * dt;
      if (sig_next == 0.f) { ph.x_state[e] = nx; }
      else { ph.dbuf[e] = dd; ph.x_aux[e] = nx; }
    } break;
    case HEAD_HEUN2: {   // sig = sigma_{i+1}; x_in = x2; x_state still holds x
      float s0 = ph.sigmas[ph.step], dt = sig - s0;
      float d2 = (xin - D) / sig;
      float dp = (__ldcg(ph.dbuf + e) + d2) / 2.0f;
      nx = __ldcg(ph.x_state + e) + dp * dt;
      ph.x_state[e] = nx;
    } break;
    case HEAD_DPMPP2M: {
      float t = -logf(sig), tn = -logf(sig_next), h = tn - t;
      float ratio = expf(-tn) / expf(-t), em = expm1f(-h);
      float den = D;
      if (ph.step > 0 && sig_next != 0.f) {
        float h_last = t - (-logf(ph.sigmas[ph.step - 1]));
        float r = h_last / h;
        den = (1.0f + 1.0f / (2.0f * r)) * D - (1.0f / (2.0f * r)) * __ldcg(ph.dbuf + e);
      }
      nx = ratio * xin - em * den;
      ph.x_state[e] = nx;
      ph.dbuf[e] = D;
    } break;
    default: break;
  }
  return nx;
}

// ------------------------------------------------------------------------------------------ ROW phase: one warp per token row
// Parameter vectors of a row phase.  With a batch-shared sigma (sampling) they are staged once per phase in shared memory by all
// workers (stage_row_params); with per-sample AdaLN rows (denoise API) shift / scale / gate stay in global memory.
struct RowPtrs {
  const float *ln_w, *ln_b, *shift, *scale, *pbias, *rgate, *dln_w, *dln_b, *ap_w, *ae_w, *ae_b;
};

__device__ __forceinline__ void stage_row_params(const Phase& ph, const FusedParams& P, float* sm, RowPtrs& rp, int wt) {
  const int d = P.d;
  rp = RowPtrs{ph.ln ? ph.ln_w : nullptr, ph.ln ? ph.ln_b : nullptr, ph.ln ? ph.shift : nullptr, ph.ln ? ph.scale : nullptr,
               ph.n_part > 0 ? ph.pbias : nullptr, ph.n_part > 0 ? ph.rgate : nullptr,
               ph.head_mode >= 0 ? ph.dln_w : nullptr, ph.head_mode >= 0 ? ph.dln_b : nullptr, ph.head_mode >= 0 ? ph.ap_w : nullptr,
               ph.embed ? ph.ae_w : nullptr, ph.embed ? ph.ae_b : nullptr};
  int off = 0;
  auto put = [&](const float*& p, int n, bool shared_by_batch) {
    if (!p || !shared_by_batch) return;
    for (int e = wt * 4; e < n; e += FD_WORKERS * 4) *reinterpret_cast<float4*>(sm + off + e) = *reinterpret_cast<const float4*>(p + e);
    p = sm + off;
    off += n;
  };
  put(rp.ln_w, d, true); put(rp.ln_b, d, true); put(rp.shift, d, ph.mod_stride == 0); put(rp.scale, d, ph.mod_stride == 0);
  put(rp.pbias, d, true); put(rp.rgate, d, ph.rgate_stride == 0);
  put(rp.dln_w, d, true); put(rp.dln_b, d, true); put(rp.ap_w, P.A * d, (P.A * d) % 4 == 0);
  put(rp.ae_w, P.A * d, (P.A * d) % 4 == 0); put(rp.ae_b, d, true);
}

// LayerNorm (+ AdaLN modulate) of the row held in v -> split-bf16 operand (ln_mod_kernel); parameters already in registers
template <int VPL>
__device__ __forceinline__ void ln_apply(const float4 (&v)[VPL], const float4 (&w)[VPL], const float4 (&bb)[VPL], bool has_b,
                                         const float4 (&sh)[VPL], const float4 (&sc)[VPL], bool has_mod, __nv_bfloat16* po, int lo_off, int lane) {
  constexpr int d = VPL * 128;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)d;
  float qv = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    qv += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(qv) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    float o[4] = {(v[i].x - mean) * rstd * w[i].x, (v[i].y - mean) * rstd * w[i].y, (v[i].z - mean) * rstd * w[i].z, (v[i].w - mean) * rstd * w[i].w};
    if (has_b) { o[0] += bb[i].x; o[1] += bb[i].y; o[2] += bb[i].z; o[3] += bb[i].w; }
    if (has_mod) { o[0] = sh[i].x + o[0] * sc[i].x; o[1] = sh[i].y + o[1] * sc[i].y; o[2] = sh[i].z + o[2] * sc[i].z; o[3] = sh[i].w + o[3] * sc[i].w; }
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(o[j], hi[j], lo[j]);
    *reinterpret_cast<uint2*>(po + c) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(po + lo_off + c) = *reinterpret_cast<uint2*>(lo);
  }
}

template <int VPL>
__device__ __forceinline__ void ln_load_params(const float* ln_w, const float* ln_b, const float* shift, const float* scale, int lane,
                                               float4 (&w)[VPL], float4 (&bb)[VPL], float4 (&sh)[VPL], float4 (&sc)[VPL]) {
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    w[i] = *reinterpret_cast<const float4*>(ln_w + c);
    bb[i] = ln_b ? *reinterpret_cast<const float4*>(ln_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sh[i] = shift ? *reinterpret_cast<const float4*>(shift + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sc[i] = shift ? *reinterpret_cast<const float4*>(scale + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
}

template <int VPL>
__device__ __forceinline__ void ln_store(const float4 (&v)[VPL], const float* ln_w, const float* ln_b, const float* shift, const float* scale,
                                         __nv_bfloat16* po, int lo_off, int lane) {
  float4 w[VPL], bb[VPL], sh[VPL], sc[VPL];
  ln_load_params<VPL>(ln_w, ln_b, shift, scale, lane, w, bb, sh, sc);
  ln_apply<VPL>(v, w, bb, ln_b != nullptr, sh, sc, shift != nullptr, po, lo_off, lane);
}

// Pure LayerNorm phase with batch-shared parameters: every operand of up to two rows is requested up front (one L2 round trip).
template <int VPL>
__device__ __noinline__ void ln_rows2(const Phase& ph, int rowA, int rowB, int lane) {
  constexpr int d = VPL * 128;
  float4 a[VPL], b[VPL], w[VPL], bb[VPL], sh[VPL], sc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    a[i] = ldcg4(ph.xh + (size_t)rowA * d + c);
    if (rowB >= 0) b[i] = ldcg4(ph.xh + (size_t)rowB * d + c);
  }
  ln_load_params<VPL>(ph.ln_w, ph.ln_b, ph.shift, ph.scale, lane, w, bb, sh, sc);
  ln_apply<VPL>(a, w, bb, ph.ln_b != nullptr, sh, sc, ph.shift != nullptr, ph.out16 + (size_t)rowA * ph.ldo, ph.lo_off16, lane);
  if (rowB >= 0) ln_apply<VPL>(b, w, bb, ph.ln_b != nullptr, sh, sc, ph.shift != nullptr, ph.out16 + (size_t)rowB * ph.ldo, ph.lo_off16, lane);
}

template <int VPL>
__device__ __noinline__ void row_phase_row(const Phase& ph, const FusedParams& P, const RowPtrs& rp, int row, int lane) {
  constexpr int d = VPL * 128;
  const int T = P.T, A = P.A;
  const int smp = row / T;
  float4 v[VPL];
  // ---- A: split-K partial sum + bias -> gated residual, or plain load of the residual stream
  if (ph.n_part > 0) {
    // loads in batches of three partials (all loads of a batch in flight together), summed in partial order -> deterministic
    float4 acc[VPL], res[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      res[i] = ldcg4(ph.xh + (size_t)row * d + (i * 32 + lane) * 4);
      acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int p0 = 0; p0 < ph.n_part; p0 += 3) {
      float4 t[3][VPL];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (p0 + p < ph.n_part) t[p][i] = ldcg4(ph.part + (size_t)(p0 + p) * ph.part_stride + (size_t)row * d + (i * 32 + lane) * 4);
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int i = 0; i < VPL; ++i)
          if (p0 + p < ph.n_part) { acc[i].x += t[p][i].x; acc[i].y += t[p][i].y; acc[i].z += t[p][i].z; acc[i].w += t[p][i].w; }
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (rp.pbias) { const float4 bq = *reinterpret_cast<const float4*>(rp.pbias + c); acc[i].x += bq.x; acc[i].y += bq.y; acc[i].z += bq.z; acc[i].w += bq.w; }
      if (rp.rgate) {
        const float4 g = *reinterpret_cast<const float4*>(rp.rgate + (size_t)smp * ph.rgate_stride + c);
        v[i] = make_float4(res[i].x + g.x * acc[i].x, res[i].y + g.y * acc[i].y, res[i].z + g.z * acc[i].z, res[i].w + g.w * acc[i].w);
      } else {
        v[i] = make_float4(res[i].x + acc[i].x, res[i].y + acc[i].y, res[i].z + acc[i].z, res[i].w + acc[i].w);
      }
      *reinterpret_cast<float4*>(ph.xh + (size_t)row * d + c) = v[i];
    }
  } else if (!(ph.embed && ph.head_mode < 0)) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = ldcg4(ph.xh + (size_t)row * d + (i * 32 + lane) * 4);
  }
  // ---- B: output head + sampler update (head_kernel of kernels_simt.cuh)
  float xnext = 0.f;   // lane j < A: element j of the actions the next evaluation sees
  if (ph.head_mode >= 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)d;
    float qv = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      qv += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(qv) / (float)d + 1e-5f);
    float4 u[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 w = *reinterpret_cast<const float4*>(rp.dln_w + c);
      u[i].x = (v[i].x - mean) * rstd * w.x; u[i].y = (v[i].y - mean) * rstd * w.y;
      u[i].z = (v[i].z - mean) * rstd * w.z; u[i].w = (v[i].w - mean) * rstd * w.w;
      if (rp.dln_b) {
        const float4 bb = *reinterpret_cast<const float4*>(rp.dln_b + c);
        u[i].x += bb.x; u[i].y += bb.y; u[i].z += bb.z; u[i].w += bb.w;
      }
    }
    float pj = 0.f;
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      const float* wr = rp.ap_w + (size_t)j * d;
      float p = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float4 w = *reinterpret_cast<const float4*>(wr + (i * 32 + lane) * 4);
        p += (u[i].x * w.x + u[i].y * w.y) + (u[i].z * w.z + u[i].w * w.w);
      }
      p = warp_sum(p);
      if (lane == j) pj = p;
    }
    if (lane < A) xnext = head_update(ph, ph.head_mode, (size_t)row * A + lane, row, T, pj + ph.ap_b[lane], P.sigma_data);
  }
  // ---- B': action embedding of the next evaluation (action_embed_kernel): xh = W_ae (x * c_in) + b_ae
  if (ph.embed) {
    if (ph.emb_x && lane < A) xnext = __ldcg(ph.emb_x + (size_t)row * A + lane);
    float c_in = 1.f;
    if (ph.precondition) {
      float cs, co;
      edm_scalings(ph.emb_sigma[(size_t)smp * ph.emb_sigma_stride], P.sigma_data, cs, co, c_in);
    }
    float acc[VPL][4];
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[i][t] = 0.f;
#pragma unroll 1
    for (int j = 0; j < A; ++j) {          // same accumulation order (j ascending) as action_embed_kernel
      const float xj = __shfl_sync(0xffffffffu, xnext, j) * c_in;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[i][t] = fmaf(rp.ae_w[(size_t)((i * 32 + lane) * 4 + t) * A + j], xj, acc[i][t]);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 bb = *reinterpret_cast<const float4*>(rp.ae_b + c);
      v[i] = make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
      *reinterpret_cast<float4*>(ph.xh + (size_t)row * d + c) = v[i];
    }
  }
  // ---- C: LayerNorm (+ AdaLN modulate) -> split-bf16 operand (ln_mod_kernel)
  if (ph.ln) {
    const size_t mrow = (size_t)smp * ph.mod_stride;
    ln_store<VPL>(v, rp.ln_w, rp.ln_b, rp.shift ? rp.shift + mrow : nullptr, rp.shift ? rp.scale + mrow : nullptr,
                  ph.out16 + (size_t)row * ph.ldo, ph.lo_off16, lane);
  }
}

// ------------------------------------------------------------------------------------------ ATTN phase (512 worker threads)
// Samples smp0 + {cta, cta + C, ...} of the row group, two at a time, staged in shared memory (attention_kernel of kernels_simt.cuh).
__device__ __noinline__ void attn_phase_generic(const Phase& ph, const FusedParams& P, float* scr, int smp0, int n_smp, int cta, int wt) {
  const int D = P.d, DP = D + 4, D4 = D / 4, H = P.H, hd = P.hd, Tq = ph.Tq, Tk = ph.Tk;
  const int per = (Tq + 2 * Tk) * DP;
  float* sp_base = scr + 2 * per;
  for (int base = cta; base < n_smp; base += 2 * P.C) {
    const int ns = (base + P.C < n_smp) ? 2 : 1;
    for (int e = wt; e < ns * Tq * D4; e += FD_WORKERS) {
      const int s = e / (Tq * D4), r = (e / D4) % Tq, c = (e % D4) * 4;
      const int b = smp0 + base + s * P.C;
      *reinterpret_cast<float4*>(scr + s * per + r * DP + c) = ldcg4(ph.q + (size_t)(b * Tq + r) * ph.ldq + c);
    }
    for (int e = wt; e < ns * Tk * D4; e += FD_WORKERS) {
      const int s = e / (Tk * D4), r = (e / D4) % Tk, c = (e % D4) * 4;
      const int b = smp0 + base + s * P.C;
      *reinterpret_cast<float4*>(scr + s * per + (Tq + r) * DP + c) = ldcg4(ph.k + (size_t)(b * Tk + r) * ph.ldkv + c);
      *reinterpret_cast<float4*>(scr + s * per + (Tq + Tk + r) * DP + c) = ldcg4(ph.v + (size_t)(b * Tk + r) * ph.ldkv + c);
    }
    bar_workers();
    for (int e = wt; e < ns * H * Tq * Tk; e += FD_WORKERS) {
      const int s = e / (H * Tq * Tk), h = (e / (Tq * Tk)) % H, i = (e / Tk) % Tq, j = e % Tk;
      float sc = -INFINITY;
      if (!(ph.causal && j > i)) {
        const float* qp = scr + s * per + i * DP + h * hd;
        const float* kp = scr + s * per + (Tq + j) * DP + h * hd;
        float acc = 0.f;
        for (int c = 0; c < hd; c += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qp + c), kv = *reinterpret_cast<const float4*>(kp + c);
          acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
        }
        sc = acc * ph.att_scale;
      }
      sp_base[((s * H + h) * Tq + i) * (Tk + 1) + j] = sc;
    }
    bar_workers();
    if (wt < ns * H * Tq) {
      float* rowp = sp_base + wt * (Tk + 1);
      float mx = -INFINITY;
      for (int j = 0; j < Tk; ++j) mx = fmaxf(mx, rowp[j]);
      float sum = 0.f;
      for (int j = 0; j < Tk; ++j) { float ex = expf(rowp[j] - mx); rowp[j] = ex; sum += ex; }
      const float inv = 1.0f / sum;
      for (int j = 0; j < Tk; ++j) rowp[j] *= inv;
    }
    bar_workers();
    for (int e = wt; e < ns * Tq * D4; e += FD_WORKERS) {
      const int s = e / (Tq * D4), i = (e / D4) % Tq, col = (e % D4) * 4, h = col / hd;
      const int b = smp0 + base + s * P.C;
      const float* pr = sp_base + ((s * H + h) * Tq + i) * (Tk + 1);
      const float* sv = scr + s * per + (Tq + Tk) * DP + col;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < Tk; ++j) {
        const float pj = pr[j];
        const float4 vv = *reinterpret_cast<const float4*>(sv + j * DP);
        o[0] = fmaf(pj, vv.x, o[0]); o[1] = fmaf(pj, vv.y, o[1]); o[2] = fmaf(pj, vv.z, o[2]); o[3] = fmaf(pj, vv.w, o[3]);
      }
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) split_bf16(o[t], hi[t], lo[t]);
      __nv_bfloat16* po = ph.out16 + (size_t)(b * Tq + i) * ph.ldo + col;
      *reinterpret_cast<uint2*>(po) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(po + ph.lo_off16) = *reinterpret_cast<uint2*>(lo);
    }
    bar_workers();
  }
}

// Compile-time specialisation for the shipped shapes: index arithmetic folds to constants and the q/k/v rows of both samples are
// requested in one batch (12 float4 per thread in flight) instead of one dependent load per loop iteration.
template <int D, int HD, int TQ, int TK>
__device__ __noinline__ void attn_phase_t(const Phase& ph, const FusedParams& P, float* scr, int smp0, int n_smp, int cta, int wt) {
  constexpr int H = D / HD, DP = D + 4, D4 = D / 4, RS = TQ + 2 * TK, PER = RS * DP, UNR = 6;
  float* sp_base = scr + 2 * PER;
  for (int base = cta; base < n_smp; base += 2 * P.C) {
    const int ns = (base + P.C < n_smp) ? 2 : 1;
    const int total = ns * RS * D4;
    for (int e0 = wt; e0 < total; e0 += UNR * FD_WORKERS) {
      float4 t[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int e = e0 + u * FD_WORKERS;
        if (e < total) {
          const int s = e / (RS * D4), rr = (e / D4) % RS, c = (e % D4) * 4;
          const int b = smp0 + base + s * P.C;
          const float* src = rr < TQ ? ph.q + (size_t)(b * TQ + rr) * ph.ldq + c
                           : rr < TQ + TK ? ph.k + (size_t)(b * TK + rr - TQ) * ph.ldkv + c : ph.v + (size_t)(b * TK + rr - TQ - TK) * ph.ldkv + c;
          t[u] = ldcg4(src);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int e = e0 + u * FD_WORKERS;
        if (e < total) {
          const int s = e / (RS * D4), rr = (e / D4) % RS, c = (e % D4) * 4;
          *reinterpret_cast<float4*>(scr + s * PER + rr * DP + c) = t[u];
        }
      }
    }
    bar_workers();
    for (int e = wt; e < ns * H * TQ * TK; e += FD_WORKERS) {
      const int s = e / (H * TQ * TK), h = (e / (TQ * TK)) % H, i = (e / TK) % TQ, j = e % TK;
      float sc = -INFINITY;
      if (!(ph.causal && j > i)) {
        const float* qp = scr + s * PER + i * DP + h * HD;
        const float* kp = scr + s * PER + (TQ + j) * DP + h * HD;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HD; c += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(qp + c), kv = *reinterpret_cast<const float4*>(kp + c);
          acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
        }
        sc = acc * ph.att_scale;
      }
      sp_base[((s * H + h) * TQ + i) * (TK + 1) + j] = sc;
    }
    bar_workers();
    if (wt < ns * H * TQ) {
      float* rowp = sp_base + wt * (TK + 1);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < TK; ++j) mx = fmaxf(mx, rowp[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < TK; ++j) { float ex = expf(rowp[j] - mx); rowp[j] = ex; sum += ex; }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int j = 0; j < TK; ++j) rowp[j] *= inv;
    }
    bar_workers();
    for (int e = wt; e < ns * TQ * D4; e += FD_WORKERS) {
      const int s = e / (TQ * D4), i = (e / D4) % TQ, col = (e % D4) * 4, h = col / HD;
      const int b = smp0 + base + s * P.C;
      const float* pr = sp_base + ((s * H + h) * TQ + i) * (TK + 1);
      const float* sv = scr + s * PER + (TQ + TK) * DP + col;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < TK; ++j) {
        const float pj = pr[j];
        const float4 vv = *reinterpret_cast<const float4*>(sv + j * DP);
        o[0] = fmaf(pj, vv.x, o[0]); o[1] = fmaf(pj, vv.y, o[1]); o[2] = fmaf(pj, vv.z, o[2]); o[3] = fmaf(pj, vv.w, o[3]);
      }
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) split_bf16(o[t], hi[t], lo[t]);
      __nv_bfloat16* po = ph.out16 + (size_t)(b * TQ + i) * ph.ldo + col;
      *reinterpret_cast<uint2*>(po) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(po + ph.lo_off16) = *reinterpret_cast<uint2*>(lo);
    }
    bar_workers();
  }
}

__device__ __forceinline__ void attn_phase(const Phase& ph, const FusedParams& P, float* scr, int smp0, int n_smp, int cta, int wt) {
  if (P.d == 384 && P.hd == 48 && ph.Tq == 10 && ph.Tk == 10) attn_phase_t<384, 48, 10, 10>(ph, P, scr, smp0, n_smp, cta, wt);
  else if (P.d == 384 && P.hd == 48 && ph.Tq == 10 && ph.Tk == 4) attn_phase_t<384, 48, 10, 4>(ph, P, scr, smp0, n_smp, cta, wt);
  else if (P.d == 512 && P.hd == 64 && ph.Tq == 10 && ph.Tk == 10) attn_phase_t<512, 64, 10, 10>(ph, P, scr, smp0, n_smp, cta, wt);
  else if (P.d == 512 && P.hd == 64 && ph.Tq == 10 && ph.Tk == 3) attn_phase_t<512, 64, 10, 3>(ph, P, scr, smp0, n_smp, cta, wt);
  else attn_phase_generic(ph, P, scr, smp0, n_smp, cta, wt);
}

// ------------------------------------------------------------------------------------------ the kernel
template <int VPL>
__global__ void __launch_bounds__(FD_THREADS, 1) fused_decoder_kernel(const FusedParams P) {
  extern __shared__ __align__(1024) uint8_t fd_smem[];
  const uint32_t sbase = smem_u32(fd_smem);
  const uint32_t bar0 = sbase + BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (2 + s); };
  auto accf_bar = [&](int b) { return bar0 + 8u * (4 + b); };
  auto acce_bar = [&](int b) { return bar0 + 8u * (6 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fd_smem + BAR_OFF + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = P.C, cta = blockIdx.x % C, group = blockIdx.x / C, n_groups = gridDim.x / C;
  const int RG = P.SPG * P.T, M = P.B * P.T;

  if ((sbase & 1023u) != 0) __trap();      // SWIZZLE_128B tiles need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(bar0 + 8u * i, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait(KT_FUSED);

  const int passes = P.passes;
  if (warp >= 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
  if (warp == 16) {
    // =================================================================== TMA producer
    int uses[2] = {0, 0};
    for (int rg = group; rg < P.n_rowgroups; rg += n_groups) {
      const int* prow = P.progress + (size_t)rg * 32;
      const int row0 = rg * RG;
      bool w0_done = false;                       // first weight tile of the current GEMM already requested (look-ahead)
      GemmDesc g = P.gemms[0];
      for (int gi = 0; gi < P.n_gemms; ++gi) {
        const CUtensorMap* ma = P.maps + g.a_map;
        const CUtensorMap* mw = P.maps + g.w_map;
        const int kidx = cta % g.cta_mod, nidx = cta / g.cta_mod;
        const int a_col = g.a_col_base + kidx * g.a_col_s1;
        const int w_row = g.w_row_base + kidx * g.w_row_s1 + nidx * g.w_row_s2, w_col = g.w_col_base + kidx * g.w_col_s1;
        const uint32_t bytes = (passes == 3 ? 2u : 1u) * (uint32_t)(A_TILE + g.bn * 128);
        const int nkb = g.nkb, a_lo = g.a_lo_off + a_col, w_lo = g.w_lo_off + w_col, dep = g.dep;
        GemmDesc nx = g;
        if (gi + 1 < P.n_gemms) nx = P.gemms[gi + 1];      // requested before the group wait: the L2 round trip overlaps it
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb & 1;
          const uint32_t st = sbase + s * STAGE;
          if (!(kb == 0 && w0_done)) {
            mbar_wait(empty_bar(s), (uint32_t)((uses[s] & 1) ^ 1));
            if (lane == 0) {
              mbar_expect_tx(full_bar(s), bytes);
              tma_load_2d(st + 2 * A_TILE, mw, full_bar(s), w_col + kb * BK, w_row);
              if (passes == 3) tma_load_2d(st + 2 * A_TILE + W_TILE_MAX, mw, full_bar(s), w_lo + kb * BK, w_row);
            }
          }
          if (kb == 0) { group_wait(prow, C, dep, lane); fence_proxy_async_all(); }
          if (lane == 0) {
            tma_load_2d(st, ma, full_bar(s), a_col + kb * BK, row0);
            if (passes == 3) tma_load_2d(st + A_TILE, ma, full_bar(s), a_lo + kb * BK, row0);
          }
          uses[s]++;
          __syncwarp();
        }
        // look-ahead: the first weight tile of the next GEMM phase does not depend on anything this group computes
        w0_done = false;
        if (gi + 1 < P.n_gemms) {
          const CUtensorMap* mw2 = P.maps + nx.w_map;
          const int k2 = cta % nx.cta_mod, n2 = cta / nx.cta_mod;
          const int w_row2 = nx.w_row_base + k2 * nx.w_row_s1 + n2 * nx.w_row_s2, w_col2 = nx.w_col_base + k2 * nx.w_col_s1;
          const uint32_t bytes2 = (passes == 3 ? 2u : 1u) * (uint32_t)(A_TILE + nx.bn * 128);
          mbar_wait(empty_bar(0), (uint32_t)((uses[0] & 1) ^ 1));
          if (lane == 0) {
            mbar_expect_tx(full_bar(0), bytes2);
            tma_load_2d(sbase + 2 * A_TILE, mw2, full_bar(0), w_col2, w_row2);
            if (passes == 3) tma_load_2d(sbase + 2 * A_TILE + W_TILE_MAX, mw2, full_bar(0), nx.w_lo_off + w_col2, w_row2);
          }
          w0_done = true;
          __syncwarp();
        }
        g = nx;
      }
    }
  } else if (warp == 17) {
    // =================================================================== MMA issuer
    int fuses[2] = {0, 0}, accn[2] = {0, 0};
    for (int rg = group; rg < P.n_rowgroups; rg += n_groups) {
      for (int gi = 0; gi < P.n_gemms; ++gi) {
        const int b = P.gemms[gi].acc_buf, nkb = P.gemms[gi].nkb, bn = P.gemms[gi].bn, p = P.gemms[gi].p;
        mbar_wait(acce_bar(b), (uint32_t)((accn[b] & 1) ^ 1));     // the epilogue that last used this TMEM buffer has drained it
        tc_fence_after();
        const uint32_t idesc = make_idesc(BM, bn);
        const uint32_t tacc = tmem_base + (uint32_t)(b * ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb) {
          const int s = kb & 1;
          mbar_wait(full_bar(s), (uint32_t)(fuses[s] & 1));
          tc_fence_after();
          if (lane == 0) {
            if (kb == 0) FD_STAMP(p, 3);
            const uint32_t st = sbase + s * STAGE;
            const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_TILE);
            const uint64_t w_hi = make_smem_desc(st + 2 * A_TILE), w_lo = make_smem_desc(st + 2 * A_TILE + W_TILE_MAX);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 2) >> 4);
              if (passes == 3) {
                umma_f16(tacc, a_lo + adv, w_hi + adv, idesc, (kb | k) != 0);   // small terms first (same order as tc_gemm_kernel)
                umma_f16(tacc, a_hi + adv, w_lo + adv, idesc, 1);
                umma_f16(tacc, a_hi + adv, w_hi + adv, idesc, 1);
              } else {
                umma_f16(tacc, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
              }
            }
            umma_commit(empty_bar(s));
          }
          fuses[s]++;
          __syncwarp();
        }
        if (lane == 0) umma_commit(accf_bar(b));
        accn[b]++;
        __syncwarp();
      }
    }
  } else if (warp < 16) {
    // =================================================================== workers (16 warps)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory");
    const int ww = warp, wt = threadIdx.x;
    const int q = warp & 3, cg = ww >> 2;          // TMEM lane quarter (hardware: warp id % 4) and 16-column group
    float* stg = reinterpret_cast<float*>(fd_smem + SCR_OFF);
    float* sbias = reinterpret_cast<float*>(fd_smem + SCR_OFF + STG_BYTES);
    float* sgate = sbias + EPI_VEC;
    float* row_scr = reinterpret_cast<float*>(fd_smem + ROW_SCR_OFF);
    Phase* sph = reinterpret_cast<Phase*>(fd_smem + PH_OFF);
    int accn[2] = {0, 0};
    for (int rg = group; rg < P.n_rowgroups; rg += n_groups) {
      int* prow = P.progress + (size_t)rg * 32;
      const int row0 = rg * RG;
      const int rows_valid = (M - row0) < RG ? (M - row0) : RG;
      const int smp0 = rg * P.SPG;
      const int n_smp = (P.B - smp0) < P.SPG ? (P.B - smp0) : P.SPG;
      for (int p = 0;; ++p) {
        // warp 0 fetches the descriptor (one uint4 per lane) while it polls the group counters, then shares it through smem
        if (ww == 0) {
          uint4 dv = make_uint4(0u, 0u, 0u, 0u);
          if (lane * 16 < (int)sizeof(Phase)) dv = __ldg(reinterpret_cast<const uint4*>(P.prog + p) + lane);
          const int dep = __shfl_sync(0xffffffffu, (int)dv.y, 0);      // Phase::dep is the second int
          group_wait(prow, C, dep, lane);
          reinterpret_cast<uint4*>(sph)[lane] = dv;
        }
        bar_workers();
        const Phase& ph = *sph;
        if (ph.type == PH_END) break;
        if (wt == 0) FD_STAMP(p, 0);
        if (ph.type == PH_GEMM) {
          const int b = ph.acc_buf;
          const int kidx = cta % ph.cta_mod, nidx = cta / ph.cta_mod;
          const int w_row = ph.w_row_base + kidx * ph.w_row_s1 + nidx * ph.w_row_s2;
          const int o_col0 = ph.o_col_base + kidx * ph.o_col_s1 + nidx * ph.o_col_s2;
          float* outp = ph.out ? ph.out + (size_t)kidx * ph.out_cta_stride : nullptr;
          // epilogue operands are requested NOW, while the mainloop runs: bias / batch-shared gate slices -> shared memory,
          // residual tile (single-chunk GEMMs) -> registers
          const bool gate_shared = ph.gate && ph.gate_stride == 0;
          if (wt < ph.bn) {
            sbias[wt] = ph.bias ? ph.bias[w_row + wt] : 0.f;
            if (gate_shared) sgate[wt] = ph.gate[o_col0 + wt];
          }
          const bool pre = ph.epi == FE_RESID && ph.bn == 64;
          float4 pre_r[4];
          if (pre) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int idx = wt + it * FD_WORKERS, r = idx >> 4, c4 = (idx & 15) * 4;
              if (r < rows_valid) pre_r[it] = ldcg4(outp + (size_t)(row0 + r) * ph.ldo + o_col0 + c4);
            }
          }
          bar_workers();
          mbar_wait(accf_bar(b), (uint32_t)(accn[b] & 1));
          accn[b]++;
          tc_fence_after();
          if (wt == 0) FD_STAMP(p, 1);
          const int nch = ph.bn / 64;
          for (int ch = 0; ch < nch; ++ch) {
            {   // phase A: TMEM -> registers -> (+bias, GELU) -> staging tile; thread = accumulator row
              const int r = q * 32 + lane, col = ch * 64 + cg * 16;
              uint32_t v[16];
              tmem_ld16(tmem_base + (uint32_t)(b * ACC_COLS) + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 bb = *reinterpret_cast<const float4*>(sbias + col + j);
                o[j] = __uint_as_float(v[j]) + bb.x; o[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
                o[j + 2] = __uint_as_float(v[j + 2]) + bb.z; o[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
              }
              if (ph.epi == FE_GELU16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = gelu_erf_fast(o[j]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<float4*>(stg + r * 64 + (((cg * 4 + j) ^ (r & 15)) << 2)) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            if (ch == 0 && wt == 0) FD_STAMP(p, 4);
            tc_fence_before();
            bar_workers();
            if (ch == 0 && wt == 0) FD_STAMP(p, 5);
            if (ch == nch - 1 && wt == 0) mbar_arrive(acce_bar(b));     // every warp has finished reading this TMEM buffer
            // phase B: coalesced (16 threads x float4 = one 256-byte row segment)
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int idx = wt + it * FD_WORKERS;
              const int r = idx >> 4, s4 = idx & 15, c4 = s4 * 4;
              if (r >= rows_valid) continue;
              float4 val = *reinterpret_cast<const float4*>(stg + r * 64 + ((s4 ^ (r & 15)) << 2));
              const int ocol = o_col0 + ch * 64 + c4;
              const size_t grow = (size_t)(row0 + r);
              if (ph.epi == FE_GELU16) {
                __nv_bfloat16 hi[4], lo[4];
                split_bf16(val.x, hi[0], lo[0]); split_bf16(val.y, hi[1], lo[1]); split_bf16(val.z, hi[2], lo[2]); split_bf16(val.w, hi[3], lo[3]);
                __nv_bfloat16* po = ph.out16 + grow * ph.ldo + ocol;
                *reinterpret_cast<uint2*>(po) = *reinterpret_cast<uint2*>(hi);
                *reinterpret_cast<uint2*>(po + ph.lo_off16) = *reinterpret_cast<uint2*>(lo);
              } else {
                float* po = outp + grow * ph.ldo + ocol;
                if (ph.epi == FE_RESID) {
                  const float4 rr = pre ? pre_r[it] : ldcg4(po);
                  if (ph.gate) {
                    const float4 g = gate_shared ? *reinterpret_cast<const float4*>(sgate + ch * 64 + c4)
                                                 : *reinterpret_cast<const float4*>(ph.gate + (size_t)((row0 + r) / P.T) * ph.gate_stride + ocol);
                    val = make_float4(rr.x + g.x * val.x, rr.y + g.y * val.y, rr.z + g.z * val.z, rr.w + g.w * val.w);
                  } else {
                    val = make_float4(rr.x + val.x, rr.y + val.y, rr.z + val.z, rr.w + val.w);
                  }
                }
                *reinterpret_cast<float4*>(po) = val;
              }
            }
            if (ch == 0 && wt == 0) FD_STAMP(p, 6);
            if (ch + 1 < nch) bar_workers();     // staging tile reusable (the publish barrier covers the last chunk)
            if (ch == 0 && wt == 0) FD_STAMP(p, 7);
          }
        } else if (ph.type == PH_ATTN) {
          attn_phase(ph, P, row_scr, smp0, n_smp, cta, wt);
        } else if (ph.type == PH_ROW) {
          // rows of samples cta, cta + C, ... ; one warp per row
          const int my_smp = n_smp > cta ? (n_smp - cta + C - 1) / C : 0;
          const int nrows = my_smp * P.T;
          auto rowof = [&](int i) { return (smp0 + cta + (i / P.T) * C) * P.T + i % P.T; };
          if (ph.n_part == 0 && ph.head_mode < 0 && !ph.embed && ph.ln && ph.mod_stride == 0) {
            for (int i = ww; i < nrows; i += 32) ln_rows2<VPL>(ph, rowof(i), i + 16 < nrows ? rowof(i + 16) : -1, lane);
          } else {
            RowPtrs rp;
            stage_row_params(ph, P, row_scr, rp, wt);
            if (wt == 0) FD_STAMP(p, 4);
            bar_workers();
            if (wt == 0) FD_STAMP(p, 5);
            for (int i = ww; i < nrows; i += 16) {
              row_phase_row<VPL>(ph, P, rp, rowof(i), lane);
              if (wt == 0 && i == ww) FD_STAMP(p, 6);
            }
          }
        }
        // publish: my share of phase p is in L2 (also for async-proxy readers)
        // (CTA barrier orders every worker's stores before the leader's cumulative gpu-scope release; one thread fences)
        if (wt == 0 && ph.type != PH_GEMM) FD_STAMP(p, 1);
        bar_workers();
        if (wt == 0) { fence_proxy_async_all(); st_release_gpu(prow + cta, p + 1); FD_STAMP(p, 2); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

inline const char* configure_fused() {
  cudaError_t e = cudaFuncSetAttribute(fused_decoder_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fused_decoder_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}}  // namespace mdt::fd
