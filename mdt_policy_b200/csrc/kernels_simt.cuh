// CUDA-core (fp32) kernels of the MDT denoising path: LayerNorm(+AdaLN modulate), tiny-sequence
// attention, sigma embedding, action embedding, output head fused with the EDM preconditioner and
// the sampler update, and an exact-fp32 tiled GEMM with fused epilogues.
//
// Math follows SURVEY.md Appendix B; reference lines are cited per kernel.  No fast-math: the
// sampler relies on IEEE inf arithmetic (log(0) = -inf on the last DDIM step).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>

namespace mdt {

// Host: launch with the programmatic-stream-serialization attribute (PDL).  MDTB200_PDL=0 disables it.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MDTB200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Programmatic dependent launch (PDL): every kernel first lets its dependents start launching (their prologue
// overlaps our tail), then waits until all prerequisite grids have completed and flushed before touching memory.
// Optional in-graph kernel timeline (tools/ktrace.py, mdtb200_debug_ktrace): when armed, thread 0 of every CTA appends
// {globaltimer, tag | event | sm id | CTAs in grid | linear block id} after its dependency wait (event 0) and, for the
// main kernels, when it finishes (event 1).  One predictable branch on a __device__ pointer when disarmed.
__device__ unsigned long long* g_ktrace = nullptr;
__device__ unsigned int g_ktrace_n = 0;
__device__ unsigned int g_ktrace_cap = 0;
enum KTag : int { KT_OTHER = 0, KT_GEMM = 1, KT_LN = 2, KT_ATTN = 3, KT_HEAD = 4, KT_EMBED = 5, KT_CROSS = 6, KT_PACK = 7, KT_SKINNY = 8, KT_FUSED = 9 };
__device__ __forceinline__ void ktrace(int tag, int event) {
  if (g_ktrace != nullptr && threadIdx.x == 0) {
    unsigned long long t; unsigned int sm;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
    const unsigned int i = atomicAdd(&g_ktrace_n, 1u);
    if (i < g_ktrace_cap) {
      const unsigned long long nb = (unsigned long long)gridDim.x * gridDim.y, bi = (unsigned long long)blockIdx.y * gridDim.x + blockIdx.x;
      g_ktrace[2 * (size_t)i] = t;
      g_ktrace[2 * (size_t)i + 1] = ((unsigned long long)tag << 56) | ((unsigned long long)event << 52) | ((unsigned long long)sm << 40) | (nb << 20) | bi;
    }
  }
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait(int tag = KT_OTHER) { asm volatile("griddepcontrol.wait;" ::: "memory"); ktrace(tag, 0); }
__device__ __forceinline__ void pdl_enter(int tag = KT_OTHER) { pdl_trigger(); pdl_wait(tag); }
// late = 1: the dependents are released by a later pdl_trigger() in the kernel body instead of at entry.  A dependent tcgen05 GEMM CTA
// owns a whole SM from the moment it is resident; released at entry it idles through this kernel's own dependency wait and body, which
// costs SM time once several sub-batch chains compete for the SMs (it only needs ~1.5 us of lead for its prologue).
__device__ __forceinline__ void pdl_enter_mode(int tag, int late) { if (!late) pdl_trigger(); pdl_wait(tag); }

// ------------------------------------------------------------------------------------------
// activations (exact variants, matching ATen)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu_erf (training: activation backward, also fused into the tensor-core GEMM epilogue)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}
// Same function with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. fp32 rounding level): one exp, one
// division and five FMAs instead of libm's branchy erff -- used in the tensor-core GEMM epilogue where GELU is the
// largest non-MMA cost.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));     // MUFU.RCP (2 ulp) and ex2.approx: far below the
  float poly = fmaf(1.061405429f, t, -1.453152027f);              // 2^-17 operand rounding of the bf16x3 GEMM that consumes it
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float mish(float x) {
  float sp = x > 20.0f ? x : log1pf(expf(x));   // softplus with ATen's threshold
  return x * tanhf(sp);
}

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

// split an fp32 value into bf16 hi + bf16 lo (round-to-nearest both): x ~= hi + lo, |err| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ------------------------------------------------------------------------------------------
// Exact fp32 GEMM:  C[M,N] = epi( A[M,K] . W[N,K]^T + bias[N] )
// 128x64x16 tiles, 256 threads, 8x4 micro-tile per thread, register-prefetched double buffering.
enum Epi : int { EPI_NONE = 0, EPI_GELU = 1, EPI_MISH = 2, EPI_SILU = 3, EPI_RES = 4, EPI_RES_GATE = 5,
                 EPI_GELU16 = 6,     // tcgen05 GEMM only (training): fp32 output = pre-activation, split-bf16 output = GELU of it
                 EPI_GELUBWD16 = 7 }; // tcgen05 GEMM only (training): split-bf16 output = acc * GELU'(R), column partial sums of it

struct GemmArgs {
  const float* A; int lda;
  const float* W;            // (N, K) row-major == nn.Linear.weight
  const float* bias;         // (N) or nullptr
  float* C; int ldc;
  const float* R; int ldr;   // residual for EPI_RES / EPI_RES_GATE (may alias C)
  const float* gate;         // EPI_RES_GATE: gate[(m / rows_per_group) * gate_stride + n]
  int gate_stride; int rows_per_group;
  int M, N, K;
  int gi, go, goff;          // output-row remap: row = (m / gi) * go + goff + m % gi   (gi == 0: identity)
  // optional split-bf16 copy of the result for the tensor-core path (hi at [row, n], lo at [row, K' + n])
  __nv_bfloat16* C16; int ldc16; int lo_off;
};

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16, SG_THREADS = 256;

template <int EPI>
__global__ void __launch_bounds__(SG_THREADS) sgemm_tn_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
  pdl_enter();
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;

  // global->register staging: A tile 128x16 = 512 float4 (2 per thread), W tile 64x16 = 256 float4 (1 per thread)
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  float4 ra0, ra1, rb;
  auto gload = [&](int k0) {
    int r0 = m0 + lrow, r1 = m0 + 64 + lrow;
    ra0 = r0 < g.M ? *reinterpret_cast<const float4*>(g.A + (size_t)r0 * g.lda + k0 + lk) : make_float4(0, 0, 0, 0);
    ra1 = r1 < g.M ? *reinterpret_cast<const float4*>(g.A + (size_t)r1 * g.lda + k0 + lk) : make_float4(0, 0, 0, 0);
    int rn = n0 + lrow;
    rb = rn < g.N ? *reinterpret_cast<const float4*>(g.W + (size_t)rn * g.K + k0 + lk) : make_float4(0, 0, 0, 0);
  };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y; As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
    As[buf][lk + 0][64 + lrow] = ra1.x; As[buf][lk + 1][64 + lrow] = ra1.y; As[buf][lk + 2][64 + lrow] = ra1.z; As[buf][lk + 3][64 + lrow] = ra1.w;
    Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y; Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  gload(0);
  sstore(0);
  __syncthreads();
  const int nk = g.K / SG_BK;
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue
  const int nb = n0 + tx * 4;
  if (nb >= g.N) return;
  float bj[4] = {0.f, 0.f, 0.f, 0.f};
  if (g.bias) {
    float4 bb = *reinterpret_cast<const float4*>(g.bias + nb);
    bj[0] = bb.x; bj[1] = bb.y; bj[2] = bb.z; bj[3] = bb.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
    int row = g.gi ? (m / g.gi) * g.go + g.goff + m % g.gi : m;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float t = acc[i][j] + bj[j];
      if (EPI == EPI_GELU) t = gelu_erf(t);
      if (EPI == EPI_MISH) t = mish(t);
      if (EPI == EPI_SILU) t = silu(t);
      v[j] = t;
    }
    if (EPI == EPI_RES || EPI == EPI_RES_GATE) {
      float4 r = *reinterpret_cast<const float4*>(g.R + (size_t)row * g.ldr + nb);
      if (EPI == EPI_RES_GATE) {
        float4 gt = *reinterpret_cast<const float4*>(g.gate + (size_t)(m / g.rows_per_group) * g.gate_stride + nb);
        v[0] = r.x + gt.x * v[0]; v[1] = r.y + gt.y * v[1]; v[2] = r.z + gt.z * v[2]; v[3] = r.w + gt.w * v[3];
      } else {
        v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
      }
    }
    if (g.C) *reinterpret_cast<float4*>(g.C + (size_t)row * g.ldc + nb) = make_float4(v[0], v[1], v[2], v[3]);
    if (g.C16) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
      __nv_bfloat16* ph = g.C16 + (size_t)row * g.ldc16 + nb;
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(ph + g.lo_off) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Skinny GEMM for the sigma path (M <= 16 rows: one row per sampling step): C[M,N] = epi(A[M,K] . W[N,K]^T + b).
// One warp per output column n: the W row is read once (coalesced float4), all M dot products are accumulated in
// registers against A staged in shared memory, then warp-reduced.  The tiled kernel above needs ~50 us for these shapes
// (6 CTAs, K-long serial loop); this one is bandwidth-bound on W (a few us).
constexpr int SKINNY_MAXM = 16;
struct SkinnyArgs { const float* A; const float* W; const float* bias; float* C; int M, N, K, epi; };

__global__ void __launch_bounds__(256) skinny_gemm_kernel(SkinnyArgs g) {
  extern __shared__ __align__(16) float sA[];           // [M][K]
  pdl_enter(KT_SKINNY);
  for (int e = threadIdx.x * 4; e < g.M * g.K; e += blockDim.x * 4) *reinterpret_cast<float4*>(sA + e) = *reinterpret_cast<const float4*>(g.A + e);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= g.N) return;
  float acc[SKINNY_MAXM];
#pragma unroll
  for (int m = 0; m < SKINNY_MAXM; ++m) acc[m] = 0.f;
  const float* wr = g.W + (size_t)n * g.K;
  for (int k = lane * 4; k < g.K; k += 128) {
    const float4 w = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
    for (int m = 0; m < SKINNY_MAXM; ++m) {
      if (m < g.M) {
        const float4 a = *reinterpret_cast<const float4*>(sA + m * g.K + k);
        acc[m] = fmaf(a.x, w.x, acc[m]); acc[m] = fmaf(a.y, w.y, acc[m]); acc[m] = fmaf(a.z, w.z, acc[m]); acc[m] = fmaf(a.w, w.w, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < SKINNY_MAXM; ++m) {
    if (m < g.M) {
      float v = warp_sum(acc[m]);
      if (lane == 0) {
        v += g.bias ? g.bias[n] : 0.f;
        if (g.epi == EPI_GELU) v = gelu_erf(v);
        if (g.epi == EPI_MISH) v = mish(v);
        if (g.epi == EPI_SILU) v = silu(v);
        g.C[(size_t)m * g.N + n] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm over d (eps 1e-5, transformer_blocks.py:29-38) optionally followed by the AdaLN
// modulate  shift + LN(x) * scale  (transformer_blocks.py:262-263).  One warp per row.
// shift/scale row = (m / rows_per_group) * mod_stride  (mod_stride = 0: one sigma for the batch).
// Writes fp32 `out` and/or the split-bf16 operand copy `out16` ([row, 0..d) hi, [row, lo_off..) lo).
struct LnArgs {
  const float* x; float* out; __nv_bfloat16* out16; int ld16; int lo_off;
  const float* w; const float* b;            // LN affine (b may be null)
  const float* shift; const float* scale;    // may be null (plain LN)
  int mod_stride; int rows_per_group;
  int M, d;
  int late;                                  // PDL: release the dependents after the own dependency wait (pdl_enter_mode)
};

template <int VPL>   // float4 vectors per lane: d = 128 * VPL
__global__ void __launch_bounds__(256) ln_mod_kernel(LnArgs a) {
  pdl_enter_mode(KT_LN, a.late);
  if (a.late) pdl_trigger();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const float* xr = a.x + (size_t)warp * a.d;
  const size_t mrow = (size_t)(warp / a.rows_per_group) * a.mod_stride;
  // every global operand is requested up front (one L2 round trip instead of two dependent ones)
  float4 v[VPL], w[VPL], bb[VPL], sh[VPL], sc[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    v[i] = *reinterpret_cast<const float4*>(xr + c);
    w[i] = *reinterpret_cast<const float4*>(a.w + c);
    bb[i] = a.b ? *reinterpret_cast<const float4*>(a.b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sh[i] = a.shift ? *reinterpret_cast<const float4*>(a.shift + mrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sc[i] = a.shift ? *reinterpret_cast<const float4*>(a.scale + mrow + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)a.d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)a.d + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    float o[4] = {(v[i].x - mean) * rstd * w[i].x, (v[i].y - mean) * rstd * w[i].y, (v[i].z - mean) * rstd * w[i].z, (v[i].w - mean) * rstd * w[i].w};
    if (a.b) { o[0] += bb[i].x; o[1] += bb[i].y; o[2] += bb[i].z; o[3] += bb[i].w; }
    if (a.shift) {
      o[0] = sh[i].x + o[0] * sc[i].x; o[1] = sh[i].y + o[1] * sc[i].y; o[2] = sh[i].z + o[2] * sc[i].z; o[3] = sh[i].w + o[3] * sc[i].w;
    }
    if (a.out) *reinterpret_cast<float4*>(a.out + (size_t)warp * a.d + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.out16) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(o[j], hi[j], lo[j]);
      __nv_bfloat16* ph = a.out16 + (size_t)warp * a.ld16 + c;
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(ph + a.lo_off) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Attention for tiny sequences (Tq, Tk <= 16), one warp per (sample, head).
// transformer_blocks.py:119-158: softmax(q k^T / sqrt(hd) + mask) v, mask[i][j] = (j <= i) when
// causal -- top-left aligned also for the non-square cross-attention (Tq=10, Tk=4).
// q rows: q + (b*Tq + i) * ldq + head*hd ; k/v rows: (b*Tk + j) * ldkv.
struct AttnArgs {
  const float* q; int ldq;
  const float* k; const float* v; int ldkv;
  float* y; int ldy;                       // fp32 out (may be null)
  __nv_bfloat16* y16; int ld16; int lo_off;  // split-bf16 out (may be null)
  int B, H, hd, Tq, Tk, causal;
  float scale;
  float p_drop; unsigned long long seed;      // training only: dropout on the attention probabilities (generic kernel)
  int late;                                   // PDL: release the dependents after the scores (specialised kernel; pdl_enter_mode)
};

// counter-based uniform in [0,1): splitmix64 of (seed, element index).  Forward and backward regenerate the same mask.
// g_seed_epoch (optional, mdtb200_op_set_seed_epoch): a device counter mixed into every seed -- the host seeds are frozen when a
// training step is captured into a CUDA graph, the counter (incremented by a captured op) gives every replay fresh masks.
__device__ const unsigned long long* g_seed_epoch = nullptr;
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  if (g_seed_epoch != nullptr) seed += *g_seed_epoch * 0xD1342543DE82EF95ull;
  unsigned long long z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(unsigned)(z >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float dropout_scale(unsigned long long seed, unsigned long long idx, float p) {
  return hash_uniform(seed, idx) >= p ? 1.0f / (1.0f - p) : 0.0f;
}

constexpr int ATT_THREADS = 256;
constexpr int ATT_MAXT = 16, ATT_MAXHD = 64;

// One CTA per sample: q/k/v rows of all heads are staged in shared memory with coalesced float4 loads
// (row stride d + 4 floats keeps the per-key float4 reads of the score phase on distinct banks), then
//   phase 1: all H*Tq*Tk scores (one dot product of length hd per thread, causal entries = -inf)
//   phase 2: row softmax (H*Tq rows)          phase 3: P.V for the Tq*d outputs, written coalesced.
inline size_t attention_smem_bytes(int d, int H, int Tq, int Tk) {
  return ((size_t)(Tq + 2 * Tk) * (d + 4) + (size_t)H * Tq * (Tk + 1)) * sizeof(float);
}

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(AttnArgs a) {
  extern __shared__ __align__(16) float att_smem[];
  pdl_enter(KT_ATTN);
  const int D = a.H * a.hd, DP = D + 4, Tq = a.Tq, Tk = a.Tk, hd = a.hd;
  float* sq = att_smem;
  float* sk = sq + Tq * DP;
  float* sv = sk + Tk * DP;
  float* sp = sv + Tk * DP;            // [H][Tq][Tk+1]
  const int b = blockIdx.x, tid = threadIdx.x;
  const int D4 = D / 4;
  for (int e = tid; e < Tq * D4; e += ATT_THREADS) {
    int r = e / D4, c = (e % D4) * 4;
    *reinterpret_cast<float4*>(sq + r * DP + c) = *reinterpret_cast<const float4*>(a.q + (size_t)(b * Tq + r) * a.ldq + c);
  }
  for (int e = tid; e < Tk * D4; e += ATT_THREADS) {
    int r = e / D4, c = (e % D4) * 4;
    *reinterpret_cast<float4*>(sk + r * DP + c) = *reinterpret_cast<const float4*>(a.k + (size_t)(b * Tk + r) * a.ldkv + c);
    *reinterpret_cast<float4*>(sv + r * DP + c) = *reinterpret_cast<const float4*>(a.v + (size_t)(b * Tk + r) * a.ldkv + c);
  }
  __syncthreads();
  for (int e = tid; e < a.H * Tq * Tk; e += ATT_THREADS) {
    const int h = e / (Tq * Tk), i = (e / Tk) % Tq, j = e % Tk;
    float s = -INFINITY;
    if (!(a.causal && j > i)) {
      const float* qp = sq + i * DP + h * hd;
      const float* kp = sk + j * DP + h * hd;
      float acc = 0.f;
      for (int c = 0; c < hd; c += 4) {
        float4 qv = *reinterpret_cast<const float4*>(qp + c), kv = *reinterpret_cast<const float4*>(kp + c);
        acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
      }
      s = acc * a.scale;
    }
    sp[(h * Tq + i) * (Tk + 1) + j] = s;
  }
  __syncthreads();
  if (tid < a.H * Tq) {
    float* row = sp + tid * (Tk + 1);
    float mx = -INFINITY;
    for (int j = 0; j < Tk; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < Tk; ++j) { float ex = expf(row[j] - mx); row[j] = ex; sum += ex; }
    const float inv = 1.0f / sum;
    for (int j = 0; j < Tk; ++j) row[j] *= inv;
    if (a.p_drop > 0.f) {     // tid = h * Tq + i
      const unsigned long long base = ((unsigned long long)b * a.H * Tq + tid) * Tk;
      for (int j = 0; j < Tk; ++j) row[j] *= dropout_scale(a.seed, base + j, a.p_drop);
    }
  }
  __syncthreads();
  for (int e = tid; e < Tq * D; e += ATT_THREADS) {
    const int i = e / D, col = e % D, h = col / hd;
    const float* pr = sp + (h * Tq + i) * (Tk + 1);
    float o = 0.f;
    for (int j = 0; j < Tk; ++j) o = fmaf(pr[j], sv[j * DP + col], o);
    const size_t row = (size_t)(b * Tq + i);
    if (a.y) a.y[row * a.ldy + col] = o;
    if (a.y16) {
      __nv_bfloat16 hi, lo;
      split_bf16(o, hi, lo);
      a.y16[row * a.ld16 + col] = hi;
      a.y16[row * a.ld16 + a.lo_off + col] = lo;
    }
  }
}

// Compile-time specialisation of the same algorithm for the shipped shapes (all index arithmetic folds to
// constants, loops unroll).  One CTA handles HC heads of one sample: grid = (B, H / HC), 128 threads.
template <int HD, int TQ, int TK, int CAUSAL, int HC>
__global__ void __launch_bounds__(128) attention_fixed_kernel(AttnArgs a) {
  constexpr int DC = HC * HD, DP = DC + 4, D4 = DC / 4, NT = 128;
  __shared__ __align__(16) float sq[TQ * DP];
  __shared__ __align__(16) float sk[TK * DP];
  __shared__ __align__(16) float sv[TK * DP];
  __shared__ float sp[HC * TQ * (TK + 1)];
  pdl_enter_mode(KT_ATTN, a.late);
  const int b = blockIdx.x, c0 = blockIdx.y * DC, tid = threadIdx.x;
#pragma unroll
  for (int e = tid; e < TQ * D4; e += NT) {
    const int r = e / D4, c = (e % D4) * 4;
    *reinterpret_cast<float4*>(sq + r * DP + c) = *reinterpret_cast<const float4*>(a.q + (size_t)(b * TQ + r) * a.ldq + c0 + c);
  }
#pragma unroll
  for (int e = tid; e < TK * D4; e += NT) {
    const int r = e / D4, c = (e % D4) * 4;
    *reinterpret_cast<float4*>(sk + r * DP + c) = *reinterpret_cast<const float4*>(a.k + (size_t)(b * TK + r) * a.ldkv + c0 + c);
    *reinterpret_cast<float4*>(sv + r * DP + c) = *reinterpret_cast<const float4*>(a.v + (size_t)(b * TK + r) * a.ldkv + c0 + c);
  }
  __syncthreads();
#pragma unroll
  for (int e = tid; e < HC * TQ * TK; e += NT) {
    const int h = e / (TQ * TK), i = (e / TK) % TQ, j = e % TK;
    float s = -INFINITY;
    if (!(CAUSAL && j > i)) {
      const float* qp = sq + i * DP + h * HD;
      const float* kp = sk + j * DP + h * HD;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        float4 qv = *reinterpret_cast<const float4*>(qp + c), kv = *reinterpret_cast<const float4*>(kp + c);
        acc = fmaf(qv.x, kv.x, acc); acc = fmaf(qv.y, kv.y, acc); acc = fmaf(qv.z, kv.z, acc); acc = fmaf(qv.w, kv.w, acc);
      }
      s = acc * a.scale;
    }
    sp[(h * TQ + i) * (TK + 1) + j] = s;
  }
  __syncthreads();
  if (a.late) pdl_trigger();
  if (tid < HC * TQ) {
    float* row = sp + tid * (TK + 1);
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TK; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < TK; ++j) { float ex = expf(row[j] - mx); row[j] = ex; sum += ex; }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < TK; ++j) row[j] *= inv;
    if (a.p_drop > 0.f) {     // training: same mask stream as attention_kernel / attention_bwd_kernel, index ((b H + h) Tq + i) Tk + j
      const int hg = blockIdx.y * HC + tid / TQ, i = tid % TQ;
      const unsigned long long base = (((unsigned long long)b * a.H + hg) * TQ + i) * TK;
#pragma unroll
      for (int j = 0; j < TK; ++j) row[j] *= dropout_scale(a.seed, base + j, a.p_drop);
    }
  }
  __syncthreads();
  // P.V: each thread produces 4 consecutive output columns of one query row
#pragma unroll
  for (int e = tid; e < TQ * D4; e += NT) {
    const int i = e / D4, col = (e % D4) * 4, h = col / HD;
    const float* pr = sp + (h * TQ + i) * (TK + 1);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      const float pj = pr[j];
      const float4 vv = *reinterpret_cast<const float4*>(sv + j * DP + col);
      o[0] = fmaf(pj, vv.x, o[0]); o[1] = fmaf(pj, vv.y, o[1]); o[2] = fmaf(pj, vv.z, o[2]); o[3] = fmaf(pj, vv.w, o[3]);
    }
    const size_t row = (size_t)(b * TQ + i);
    if (a.y) *reinterpret_cast<float4*>(a.y + row * a.ldy + c0 + col) = make_float4(o[0], o[1], o[2], o[3]);
    if (a.y16) {
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) split_bf16(o[t], hi[t], lo[t]);
      __nv_bfloat16* ph = a.y16 + row * a.ld16 + c0 + col;
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(ph + a.lo_off) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Algebraic cross-attention (decoder, transformer_blocks.py:300-304 with Attention.forward :119-158).  The context K/V are
// constant over a sampling call, so per sample b, head h and context token j
//     score[i,h,j] = (LN3(x_i) Wq_h^T + bq_h) . k_hj / sqrt(hd) = LN3(x_i) . G_hj + c_hj ,   G_hj = Wq_h^T k_hj / sqrt(hd)  (d floats)
//     x_i += c_proj(sum_j p_ihj v_hj) = sum_{h,j} p_ihj U_hj + b_co ,                         U_hj = Wco[:, head h] v_hj     (d floats)
// G and U are computed ONCE per call by two head-grouped tensor-core GEMMs per layer (K = head_dim padded to 64); the per-step
// LN3 -> query GEMM -> attention -> c_proj GEMM -> LN2 chain (5 kernels) becomes ONE kernel, cross_row_kernel.

// commit time: stacked per-head weight operands (split bf16, K padded to 64):
//   wg[(h*d + m), c] = Wq[h*hd + c, m] / sqrt(hd)        wu[(h*d + n), c] = Wco[n, h*hd + c]          (c < hd, zero beyond)
__global__ void pack_cross_weights_kernel(const float* __restrict__ Wq, const float* __restrict__ Wco, __nv_bfloat16* __restrict__ wg,
                                          __nv_bfloat16* __restrict__ wu, int d, int H, int hd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * d * 64) return;
  const int c = idx % 64, m = (idx / 64) % d, h = idx / (64 * d);
  float g = 0.f, u = 0.f;
  if (c < hd) {
    g = Wq[(size_t)(h * hd + c) * d + m] * (1.0f / sqrtf((float)hd));
    u = Wco[(size_t)m * d + h * hd + c];
  }
  __nv_bfloat16 hi, lo;
  const size_t o = (size_t)(h * d + m) * 128 + c;
  split_bf16(g, hi, lo); wg[o] = hi; wg[o + 64] = lo;
  split_bf16(u, hi, lo); wu[o] = hi; wu[o + 64] = lo;
}

// per call: head-major split-bf16 operands of the context keys / values of every decoder layer, and the score constants
//   ka[l][(h*mcp + r), c] = k_l[r, h*hd + c]   va likewise   (r = b*Tc + j < Mc, zero padding to mcp rows and 64 columns)
//   ctab[l][b][h*Tc + j] = bq_l[h*hd : (h+1)*hd] . k_l[b*Tc + j, h*hd : ...] / sqrt(hd)
struct CrossPackArgs {
  const float* kv; int ldkv;          // (Mc, L*2d): layer l keys at column l*2d, values at l*2d + d
  const float* bq_all;                // (L, d): cross-attention query biases of every decoder layer
  __nv_bfloat16 *ka, *va; size_t layer_stride16;   // elements between layers
  float* ctab; size_t ctab_layer_stride;
  int Mc, mcp, Tc, H, hd, d, L;
};
__global__ void __launch_bounds__(256) pack_cross_operands_kernel(CrossPackArgs a) {
  pdl_enter(KT_PACK);
  const int l = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (h, r, c): c < 64
  if (idx < a.H * a.mcp * 64) {
    const int c = idx % 64, r = (idx / 64) % a.mcp, h = idx / (64 * a.mcp);
    float kx = 0.f, vx = 0.f;
    if (r < a.Mc && c < a.hd) {
      const float* row = a.kv + (size_t)r * a.ldkv + (size_t)l * 2 * a.d + h * a.hd + c;
      kx = row[0]; vx = row[a.d];
    }
    __nv_bfloat16 hi, lo;
    const size_t o = (size_t)l * a.layer_stride16 + (size_t)(h * a.mcp + r) * 128 + c;
    split_bf16(kx, hi, lo); a.ka[o] = hi; a.ka[o + 64] = lo;
    split_bf16(vx, hi, lo); a.va[o] = hi; a.va[o + 64] = lo;
  }
  if (idx < a.Mc * a.H) {            // one score constant per (row r = b*Tc + j, head h)
    const int h = idx % a.H, r = idx / a.H, b = r / a.Tc, j = r % a.Tc;
    const float* kr = a.kv + (size_t)r * a.ldkv + (size_t)l * 2 * a.d + h * a.hd;
    const float* bq = a.bq_all + (size_t)l * a.d + h * a.hd;
    float acc = 0.f;
    for (int c = 0; c < a.hd; ++c) acc = fmaf(bq[c], kr[c], acc);
    a.ctab[(size_t)l * a.ctab_layer_stride + (size_t)b * a.H * a.Tc + h * a.Tc + j] = acc * (1.0f / sqrtf((float)a.hd));
  }
}

struct CrossRowArgs {
  float* xh;                                   // (M, d) residual stream, updated in place
  const float* gtab; const float* utab;        // this layer / this sub-batch: rows (h*mcp + b*Tc + j), d floats each
  const float* ctab;                           // (B, H*Tc)
  int mcp;
  const float *ln3_w, *ln3_b, *bco, *ln2_w, *ln2_b, *shift, *scale; int mod_stride;
  __nv_bfloat16* a16; int ld16, lo_off;        // LN2 (+modulate) output: split-bf16 operand of c_fc
  int B, T, Tc, H, d;
  int early;                                   // 1: tables / parameters are requested BEFORE the dependency wait (see the kernel)
  int late;                                    // PDL: release the dependents after the softmax instead of at entry (pdl_enter_mode)
};
constexpr int CR_THREADS = 384;
constexpr int CR_TMAX = 12;                    // score accumulators per thread (T * 32 < CR_THREADS -> T <= 11)
// work split of the two [T x H*Tc x d] products (template parameter VPL = d / 128)
template <int VPL>
struct CrossCfg {
  static constexpr int d = VPL * 128, D4 = d / 4;
  static constexpr int KSL = d / 32;                     // scores: 32-float slices of the feature axis (12 / 16), one per warp pass
  static constexpr int NSPLIT = CR_THREADS / D4;         // output: threads per column quad (4 at d = 384, 3 at d = 512) ...
  static constexpr int NSH = VPL == 3 ? 2 : 1;           // ... = halves of the (head, token) sum ...
  static constexpr int NSR = NSPLIT / NSH;               // ... x row groups (2 / 3)
  static constexpr int RGP = VPL == 3 ? 8 : 4;           // probability slots per row group (whole float4s)
  static constexpr int RGMAX = VPL == 3 ? 6 : 4;         // rows per group the accumulators cover: NSR * RGMAX >= 11
  static constexpr int PS = NSR * RGP;                   // floats per (head, token) row of the transposed probabilities
};
inline size_t cross_row_smem_bytes(int d, int T, int Tc, int H) {
  const int HT = H * Tc;
  return ((size_t)(2 * T + HT) * (d + 4) + (size_t)((T * HT + 3) & ~3) + (size_t)HT * 16 + (size_t)(d / 32) * T * HT + (size_t)((HT + 3) & ~3) + 5 * (size_t)d) * sizeof(float);
}

// One CTA per sample.  smem: z[T][d+4] (LN3 output, later the second half-sum of the output product), x[T][d+4] (residual rows),
// tab[H*Tc][d+4] (G, later U), p[T][H*Tc] (scores), pt[H*Tc][PS] (probabilities, transposed), part[KSL][T][H*Tc] (score partials),
// ct[H*Tc] (score constants), prm[5][d].
// Both products are register-tiled so that the FMA pipe, not shared-memory bandwidth, bounds them: a scores thread owns one (head,
// token) row of G over a 32-float slice and ALL T query rows (G float4 once per T x 4 FMAs, z as warp-wide broadcasts); an output
// thread owns one column quad, 3..6 query rows and half (d = 384) or all of the (head, token) sum.
// PDL: everything except the residual rows is written once per sampling call, long before this kernel's predecessor (a tcgen05 GEMM,
// which releases its dependents only after its OWN dependency wait) could start; with a.early those operands are therefore requested
// before griddepcontrol.wait and arrive while the predecessor drains its epilogue.
template <int VPL>
__global__ void __launch_bounds__(CR_THREADS) cross_row_kernel(CrossRowArgs a) {
  extern __shared__ __align__(16) float cr_smem[];
  using C = CrossCfg<VPL>;
  constexpr int d = C::d, DP = d + 4, D4 = C::D4;
  const int T = a.T, Tc = a.Tc, H = a.H, HT = H * Tc;
  float* sz = cr_smem;
  float* sx = sz + T * DP;
  float* stab = sx + T * DP;
  float* sp = stab + HT * DP;
  float* spt = sp + ((T * HT + 3) & ~3);
  float* spart = spt + HT * 16;
  float* sct = spart + C::KSL * T * HT;
  float* sprm = sct + ((HT + 3) & ~3);   // bco | ln2_w | ln2_b | shift | scale
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!a.late) pdl_trigger();
  if (!a.early) pdl_wait(KT_CROSS);
  // ---- 1. every global operand is requested up front: G and U rows (registers), LN3 parameters (one warp per row), the small
  //         parameter vectors and score constants (remaining warps -> shared memory); then (after the wait) the T residual rows
  constexpr int TPT = 8;                // float4 of G (and of U) per thread: H*Tc*d/4 <= 8 * 384
  float4 g[TPT], u[TPT];
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    const int e = tid + t * CR_THREADS;
    if (e < HT * D4) {
      const int hj = e / D4, c = (e % D4) * 4;
      const size_t row = (size_t)(hj / Tc) * a.mcp + (size_t)b * Tc + hj % Tc;
      g[t] = *reinterpret_cast<const float4*>(a.gtab + row * d + c);
      u[t] = *reinterpret_cast<const float4*>(a.utab + row * d + c);
    }
  }
  float4 xr[VPL], w3[VPL], b3[VPL];
  if (warp < T) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      w3[i] = *reinterpret_cast<const float4*>(a.ln3_w + c);
      b3[i] = a.ln3_b ? *reinterpret_cast<const float4*>(a.ln3_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    const size_t mrow = (size_t)b * a.mod_stride;
    for (int e = (tid - T * 32) * 4; e < 5 * d; e += (CR_THREADS - T * 32) * 4) {
      const int which = e / d, c = e % d;
      const float* src = which == 0 ? a.bco : which == 1 ? a.ln2_w : which == 2 ? a.ln2_b : which == 3 ? (a.shift ? a.shift + mrow : nullptr) : (a.shift ? a.scale + mrow : nullptr);
      float4 v = src ? *reinterpret_cast<const float4*>(src + c) : (which == 4 ? make_float4(1.f, 1.f, 1.f, 1.f) : make_float4(0.f, 0.f, 0.f, 0.f));
      *reinterpret_cast<float4*>(sprm + e) = v;
    }
    for (int e = tid - T * 32; e < HT; e += CR_THREADS - T * 32) sct[e] = a.ctab[(size_t)b * HT + e];
  }
  for (int e = tid; e < HT * 16; e += CR_THREADS) spt[e] = 0.f;      // unused probability slots stay zero
  if (a.early) pdl_wait(KT_CROSS);
  if (warp < T) {
    const float* xp = a.xh + ((size_t)b * T + warp) * d;
#pragma unroll
    for (int i = 0; i < VPL; ++i) xr[i] = *reinterpret_cast<const float4*>(xp + (i * 32 + lane) * 4);
  }
  // ---- 2. G -> shared memory; LN3 of the T rows -> z, raw rows -> x
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    const int e = tid + t * CR_THREADS;
    if (e < HT * D4) *reinterpret_cast<float4*>(stab + (e / D4) * DP + (e % D4) * 4) = g[t];
  }
  if (warp < T) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (xr[i].x + xr[i].y) + (xr[i].z + xr[i].w);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float dx = xr[i].x - mean, dy = xr[i].y - mean, dz = xr[i].z - mean, dw = xr[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(sx + warp * DP + c) = xr[i];
      *reinterpret_cast<float4*>(sz + warp * DP + c) =
          make_float4((xr[i].x - mean) * rstd * w3[i].x + b3[i].x, (xr[i].y - mean) * rstd * w3[i].y + b3[i].y,
                      (xr[i].z - mean) * rstd * w3[i].z + b3[i].z, (xr[i].w - mean) * rstd * w3[i].w + b3[i].w);
    }
  }
  __syncthreads();
  // ---- 3a. score partials: lane = (head, token) row of G, warp = 32-float feature slice, T accumulators per thread
  for (int s = warp; s < C::KSL; s += CR_THREADS / 32) {
    for (int hj = lane; hj < HT; hj += 32) {
      float acc[CR_TMAX];
#pragma unroll
      for (int i = 0; i < CR_TMAX; ++i) acc[i] = 0.f;
      const float* gp = stab + hj * DP + s * 32;
      const float* zp = sz + s * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 gv = *reinterpret_cast<const float4*>(gp + c);
#pragma unroll
        for (int i = 0; i < CR_TMAX; ++i) {
          if (i < T) {
            const float4 zv = *reinterpret_cast<const float4*>(zp + i * DP + c);
            acc[i] = fmaf(zv.x, gv.x, acc[i]); acc[i] = fmaf(zv.y, gv.y, acc[i]);
            acc[i] = fmaf(zv.z, gv.z, acc[i]); acc[i] = fmaf(zv.w, gv.w, acc[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < CR_TMAX; ++i)
        if (i < T) spart[(s * T + i) * HT + hj] = acc[i];
    }
  }
  __syncthreads();
  // ---- 3b. scores = sum of the slice partials (fixed order) + constant; causal top-left mask j <= i.  U -> shared memory (G is done)
  for (int e = tid; e < T * HT; e += CR_THREADS) {
    const int i = e / HT, hj = e % HT, j = hj % Tc;
    float sc = -INFINITY;
    if (j <= i) {
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int s = 0; s < C::KSL; s += 2) { acc0 += spart[s * T * HT + e]; acc1 += spart[(s + 1) * T * HT + e]; }
      sc = (acc0 + acc1) + sct[hj];
    }
    sp[e] = sc;
  }
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    const int e = tid + t * CR_THREADS;
    if (e < HT * D4) *reinterpret_cast<float4*>(stab + (e / D4) * DP + (e % D4) * 4) = u[t];
  }
  __syncthreads();
  // ---- 4. softmax over j per (row, head) -> transposed probabilities pt[hj][row group][row in group]
  const int RG = (T + C::NSR - 1) / C::NSR;       // rows per group of the output product
  if (tid < T * H) {
    const int i = tid / H, hh = tid % H;
    const float* row = sp + i * HT + hh * Tc;
    float mx = -INFINITY;
    for (int j = 0; j < Tc; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < Tc; ++j) sum += expf(row[j] - mx);
    const float inv = 1.0f / sum;
    float* dst = spt + (size_t)(hh * Tc) * C::PS + (i / RG) * C::RGP + i % RG;
    for (int j = 0; j < Tc; ++j) dst[j * C::PS] = expf(row[j] - mx) * inv;
  }
  __syncthreads();
  if (a.late) pdl_trigger();
  // ---- 5. x_i += sum_hj p_i,hj U_hj + b_co   (kept in shared memory for the LayerNorm below, written back to the residual stream)
  {
    const int c = (tid % D4) * 4, grp = tid / D4;
    const int rg = grp % C::NSR, hs = grp / C::NSR;
    const int HH = (HT + C::NSH - 1) / C::NSH;
    const int hj_end = (hs + 1) * HH < HT ? (hs + 1) * HH : HT;
    float4 acc[C::RGMAX];
#pragma unroll
    for (int r = 0; r < C::RGMAX; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int hj = hs * HH; hj < hj_end; ++hj) {
      const float4 uv = *reinterpret_cast<const float4*>(stab + hj * DP + c);
      const float* pp = spt + hj * C::PS + rg * C::RGP;
      float pr[C::RGP];
#pragma unroll
      for (int r = 0; r < C::RGP; r += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(pp + r);
        pr[r] = pv.x; pr[r + 1] = pv.y; pr[r + 2] = pv.z; pr[r + 3] = pv.w;
      }
#pragma unroll
      for (int r = 0; r < C::RGMAX; ++r) {
        acc[r].x = fmaf(pr[r], uv.x, acc[r].x); acc[r].y = fmaf(pr[r], uv.y, acc[r].y);
        acc[r].z = fmaf(pr[r], uv.z, acc[r].z); acc[r].w = fmaf(pr[r], uv.w, acc[r].w);
      }
    }
    if (C::NSH == 2) {          // second half of the (head, token) sum -> z (free since the scores), added by the first half's thread
      if (hs == 1) {
#pragma unroll
        for (int r = 0; r < C::RGMAX; ++r) {
          const int i = rg * RG + r;
          if (r < RG && i < T) *reinterpret_cast<float4*>(sz + i * DP + c) = acc[r];
        }
      }
      __syncthreads();
    }
    if (hs == 0) {
      const float4 bc = *reinterpret_cast<const float4*>(sprm + c);
#pragma unroll
      for (int r = 0; r < C::RGMAX; ++r) {
        const int i = rg * RG + r;
        if (r < RG && i < T) {
          float4 o = acc[r];
          if (C::NSH == 2) {
            const float4 o2 = *reinterpret_cast<const float4*>(sz + i * DP + c);
            o.x += o2.x; o.y += o2.y; o.z += o2.z; o.w += o2.w;
          }
          const float4 x0 = *reinterpret_cast<const float4*>(sx + i * DP + c);
          const float4 x1 = make_float4(x0.x + (o.x + bc.x), x0.y + (o.y + bc.y), x0.z + (o.z + bc.z), x0.w + (o.w + bc.w));
          *reinterpret_cast<float4*>(sx + i * DP + c) = x1;
          *reinterpret_cast<float4*>(a.xh + ((size_t)b * T + i) * d + c) = x1;
        }
      }
    }
  }
  __syncthreads();
  // ---- 6. LN2 (+ AdaLN modulate) -> split-bf16 operand of c_fc, one warp per row
  if (warp < T) {
    float4 v[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) v[i] = *reinterpret_cast<const float4*>(sx + warp * DP + (i * 32 + lane) * 4);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
    const size_t row = (size_t)b * T + warp;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 w = *reinterpret_cast<const float4*>(sprm + d + c), bb = *reinterpret_cast<const float4*>(sprm + 2 * d + c);
      const float4 sh = *reinterpret_cast<const float4*>(sprm + 3 * d + c), sc = *reinterpret_cast<const float4*>(sprm + 4 * d + c);
      float o[4] = {(v[i].x - mean) * rstd * w.x, (v[i].y - mean) * rstd * w.y, (v[i].z - mean) * rstd * w.z, (v[i].w - mean) * rstd * w.w};
      if (a.ln2_b) { o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w; }
      if (a.shift) { o[0] = sh.x + o[0] * sc.x; o[1] = sh.y + o[1] * sc.y; o[2] = sh.z + o[2] * sc.z; o[3] = sh.w + o[3] * sc.w; }
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_bf16(o[j], hi[j], lo[j]);
      __nv_bfloat16* ph = a.a16 + row * a.ld16 + c;
      *reinterpret_cast<uint2*>(ph) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(ph + a.lo_off) = *reinterpret_cast<uint2*>(lo);
    }
  }
  ktrace(KT_CROSS, 1);
}

// ------------------------------------------------------------------------------------------
// Sinusoidal sigma embedding: pe[r, :] = [sin(e f_k), cos(e f_k)], e = log(sigma_r)/4,
// f_k = exp(-k ln(10000)/(half-1))      (mdtv_transformer.py:13-25, :238-244)
__global__ void sigma_posemb_kernel(const float* __restrict__ sigma, int R, int d, float* __restrict__ pe) {
  pdl_enter();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int half = d / 2;
  if (idx >= R * half) return;
  int r = idx / half, k = idx % half;
  float e = logf(sigma[r]) / 4.0f;
  float fstep = -(float)(9.210340371976184 / (double)(half - 1));   // -ln(10000)/(half-1), rounded once like torch
  float f = expf((float)k * fstep);
  float ang = e * f;
  pe[(size_t)r * d + k] = sinf(ang);
  pe[(size_t)r * d + half + k] = cosf(ang);
}

// ------------------------------------------------------------------------------------------
// EDM scalings (score_wrappers.py:31-43)
__device__ __forceinline__ void edm_scalings(float sigma, float sd, float& c_skip, float& c_out, float& c_in) {
  float s2 = sigma * sigma + sd * sd;
  c_skip = (sd * sd) / s2;
  c_out = sigma * sd / sqrtf(s2);
  c_in = 1.0f / sqrtf(s2);
}

// action embedding: xh[m, :] = W_ae (x[m, :] * c_in(sigma_b)) + b_ae   (score_wrappers.py:79, mdtv_transformer.py:226)
// sigma index = (m / T) * sigma_stride (0 -> one sigma for the batch).  precondition == 0: c_in = 1.
struct ActEmbArgs {
  const float* x; const float* sigma; int sigma_stride; int T; int A; int d; int M;
  const float* W; const float* b; float* xh; float sigma_data; int precondition;
};
__global__ void action_embed_kernel(ActEmbArgs a) {
  pdl_enter(KT_EMBED);
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.M * a.d) return;
  int m = idx / a.d, n = idx % a.d;
  float c_in = 1.f;
  if (a.precondition) {
    float cs, co;
    edm_scalings(a.sigma[(size_t)(m / a.T) * a.sigma_stride], a.sigma_data, cs, co, c_in);
  }
  float acc = 0.f;
  for (int j = 0; j < a.A; ++j) acc = fmaf(a.W[n * a.A + j], a.x[(size_t)m * a.A + j] * c_in, acc);
  a.xh[idx] = acc + a.b[n];
}

// ------------------------------------------------------------------------------------------
// Output head: LN_dec -> action_pred (d -> A) -> EDM precondition -> sampler update, one warp per
// token row.  Sampler formulas: gc_sampling.py:922-951 (ddim), :164-210 (euler), :256-311 (heun),
// :699-733 (dpmpp_2m); scalar coefficients are recomputed per row from the device sigma table
// with the same fp32 op sequence torch uses on 0-dim tensors.
enum HeadMode : int {
  HEAD_RAW = 0,        // out = action_pred(LN(xh))                  (forward_dec_only)
  HEAD_DENOISE = 1,    // out = raw * c_out + x * c_skip             (GCDenoiser.forward)
  HEAD_DDIM = 2,       // x <- (s'/s) x - expm1(-h) D
  HEAD_EULER = 3,      // x <- x + (x - D)/s * (s' - s)
  HEAD_HEUN1 = 4,      // d = (x - D)/s ; dbuf = d ; x2 = x + d dt   (x kept)
  HEAD_HEUN2 = 5,      // d2 = (x2 - D2)/s' ; x <- x + (dbuf + d2)/2 dt
  HEAD_DPMPP2M = 6,    // multistep with old denoised in dbuf
  HEAD_EULER_ANC = 7   // x <- x + (x - D)/s (s_down - s) + noise s_up     (sample_euler_ancestral, gc_sampling.py:213-253)
};
struct HeadArgs {
  const float* xh; const float* lnw; const float* lnb; const float* W; const float* bias;  // W (A, d)
  const float* x_in;      // actions the network was evaluated on (x, or x2 for HEUN2)
  float* x_state;         // sampler state x (updated in place for sampler modes)
  float* x_aux;           // HEUN: x2 buffer
  float* dbuf;            // HEUN: d ; DPMPP2M: old denoised
  float* out;             // RAW / DENOISE output
  const float* sigma; int sigma_stride;   // per-sample sigma (RAW/DENOISE); sampler modes use sigmas[step]
  const float* sigmas; int step; int n_steps;
  int M, d, A, T, mode; float sigma_data;
  const float* noise; const float* eta;   // EULER_ANC: this step's standard-normal draws (M x A, drawn by the caller: torch's RNG stream) and eta (device scalar)
};

template <int VPL>
__global__ void __launch_bounds__(256) head_kernel(HeadArgs a) {
  pdl_enter(KT_HEAD);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const float* xr = a.xh + (size_t)warp * a.d;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)a.d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)a.d + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 w = *reinterpret_cast<const float4*>(a.lnw + c);
    v[i].x = (v[i].x - mean) * rstd * w.x; v[i].y = (v[i].y - mean) * rstd * w.y;
    v[i].z = (v[i].z - mean) * rstd * w.z; v[i].w = (v[i].w - mean) * rstd * w.w;
    if (a.lnb) {
      float4 bb = *reinterpret_cast<const float4*>(a.lnb + c);
      v[i].x += bb.x; v[i].y += bb.y; v[i].z += bb.z; v[i].w += bb.w;
    }
  }
  // sampler scalars (uniform over the batch)
  float sig, sig_next = 0.f;
  if (a.mode <= HEAD_DENOISE) {
    sig = a.sigma ? a.sigma[(size_t)(warp / a.T) * a.sigma_stride] : 1.f;
  } else {
    sig = a.sigmas[a.mode == HEAD_HEUN2 ? a.step + 1 : a.step];
    sig_next = a.sigmas[a.step + 1];
  }
  float c_skip = 0.f, c_out = 1.f, c_in;
  if (a.mode != HEAD_RAW) edm_scalings(sig, a.sigma_data, c_skip, c_out, c_in);

  // all A dot products first (independent loads in flight together), every lane ends up with all sums ...
  float pj[1] = {0.f};
#pragma unroll 1
  for (int j = 0; j < a.A; ++j) {
    const float* wr = a.W + (size_t)j * a.d;
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      float4 w = *reinterpret_cast<const float4*>(wr + (i * 32 + lane) * 4);
      p += (v[i].x * w.x + v[i].y * w.y) + (v[i].z * w.z + v[i].w * w.w);
    }
    p = warp_sum(p);
    if (lane == j) pj[0] = p;          // lane j keeps output j
  }
  // ... then lane j finishes output j: the A sampler updates run in parallel lanes instead of serially on lane 0
  if (lane < a.A) {
    const int j = lane;
    const size_t e = (size_t)warp * a.A + j;
    const float raw = pj[0] + a.bias[j];
    if (a.mode == HEAD_RAW) { a.out[e] = raw; return; }
    const float xin = a.x_in[e];
    const float D = raw * c_out + xin * c_skip;
    switch (a.mode) {
      case HEAD_DENOISE: a.out[e] = D; break;
      case HEAD_DDIM: {
        float t = -logf(sig), tn = -logf(sig_next), h = tn - t;
        a.x_state[e] = (expf(-tn) / expf(-t)) * xin - expm1f(-h) * D;
      } break;
      case HEAD_EULER: {
        float dd = (xin - D) / sig;
        a.x_state[e] = xin + dd * (sig_next - sig);
      } break;
      case HEAD_HEUN1: {
        float dd = (xin - D) / sig, dt = sig_next - sig;
        if (sig_next == 0.f) { a.x_state[e] = xin + dd * dt; }
        else { a.dbuf[e] = dd; a.x_aux[e] = xin + dd * dt; }
      } break;
      case HEAD_HEUN2: {   // sig = sigma_{i+1}; x_in = x2; x_state still holds x
        float s0 = a.sigmas[a.step], dt = sig - s0;
        float d2 = (xin - D) / sig;
        float dp = (a.dbuf[e] + d2) / 2.0f;
        a.x_state[e] = a.x_state[e] + dp * dt;
      } break;
      case HEAD_EULER_ANC: {   // get_ancestral_step (gc_sampling.py:102-109) in torch's fp32 op order, then the Euler step to sigma_down
        const float eta = *a.eta;
        float s_down = sig_next, s_up = 0.f;
        if (eta != 0.f) {
          s_up = fminf(sig_next, eta * sqrtf(sig_next * sig_next * (sig * sig - sig_next * sig_next) / (sig * sig)));
          s_down = sqrtf(sig_next * sig_next - s_up * s_up);
        }
        float dd = (xin - D) / sig;
        float xn = xin + dd * (s_down - sig);
        if (s_down > 0.f) xn = xn + a.noise[e] * s_up;
        a.x_state[e] = xn;
      } break;
      case HEAD_DPMPP2M: {
        float t = -logf(sig), tn = -logf(sig_next), h = tn - t;
        float ratio = expf(-tn) / expf(-t), em = expm1f(-h);
        float den = D;
        if (a.step > 0 && sig_next != 0.f) {
          float h_last = t - (-logf(a.sigmas[a.step - 1]));
          float r = h_last / h;
          den = (1.0f + 1.0f / (2.0f * r)) * D - (1.0f / (2.0f * r)) * a.dbuf[e];
        }
        a.x_state[e] = ratio * xin - em * den;
        a.dbuf[e] = D;
      } break;
      default: break;
    }
  }
}

// ------------------------------------------------------------------------------------------
// MDT variant: add the learned positional embedding rows to the encoder input
// (mdt_transformer.py:318-324): token 0 += pos[0], tokens 1.. += pos[goal_seq_len] (t = 1).
__global__ void add_pos_emb_kernel(float* x, const float* pos, int B, int Tc, int d) {
  pdl_enter();
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Tc * d) return;
  int c = idx % d, t = (idx / d) % Tc;
  x[idx] += pos[(t == 0 ? 0 : 1) * d + c];
}

// fp32 -> split bf16 [rows, 2*cols] (hi | lo) conversion of a weight matrix (done once at commit)
__global__ void split_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ o, int64_t rows, int cols) {
  pdl_enter();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols; int c = (int)(idx % cols);
  __nv_bfloat16 hi, lo;
  split_bf16(w[idx], hi, lo);
  o[r * 2 * cols + c] = hi;
  o[r * 2 * cols + cols + c] = lo;
}

// pseudo-random split-bf16 operand [rows, 2*cols] for kernel timing (hash of the element index)
__global__ void fill_operand_kernel(__nv_bfloat16* __restrict__ o, int64_t rows, int cols, unsigned int seed) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const float v = 2.0f * hash_uniform(seed, (unsigned long long)idx) - 1.0f;
  int64_t r = idx / cols; int c = (int)(idx % cols);
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  o[r * 2 * cols + c] = hi;
  o[r * 2 * cols + cols + c] = lo;
}

}  // namespace mdt
