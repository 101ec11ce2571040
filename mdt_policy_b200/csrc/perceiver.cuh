// PerceiverResampler for sm_100a: the module that produces the 3 state tokens the denoiser is conditioned on
// (reference: mdt/models/networks/transformers/perceiver_resampler.py:11-163, called at mdt/models/mdtv_agent.py:392-403 on the
// (B, 1, 392, 384) Voltron token sequence; shipped config conf/model/mdtv_agent.yaml:28-32: dim 384, depth 6, 8 heads x 64).
//
// B200 design.  Per layer the reference projects K and V for all F = 392 feature tokens (2 x 395 x 384 x 512 MACs per sample),
// although only Q = 3 latent queries attend to them.  Here the queries are projected INTO feature space instead:
//     score[h,q,f] = q_hq . (Wk_h (g (.) xhat_f + b)) = (g (.) Wk_h^T q_hq) . xhat_f + q_hq . (Wk_h b)
//     out[h,q]     = sum_f a_f Wv_h (g (.) xhat_f + b) = Wv_h (g (.) z_hq) + (sum_f a_f) Wv_h b ,   z_hq = sum_f a_f xhat_f
// (g, b = norm_media affine; xhat = features normalised WITHOUT affine, computed once and shared by all layers), which needs
// 2 x (H Q) x F x d MACs per sample and layer -- 21x fewer -- and never materialises K / V.  All weight-side products run on the
// tcgen05 GEMM (gemm_tcgen05.cuh) at M = B*Q rows (the per-head value projection with head-grouped weights, N = 64), except the
// score path (q, latent keys, feature-space queries): exact fp32 on CUDA cores, because a saturated softmax amplifies score errors
// (it is tiny: B*Q rows).  The score and weighted-sum passes over xhat are SIMT kernels that stream the features once each per layer.
#pragma once
#include "gemm_tcgen05.cuh"

#include <map>
#include <string>
#include <vector>

namespace mdt { namespace pr {

// xhat[b, f, :] = normalise(x_f[b, t, n, :] + time_pos_emb[t] * mask[b, t])  (no affine), f = t * n_per_frame + n.  Warp per row.
struct PrepArgs { const float* x; const float* tpe; const float* mask; float* xhat; int B, T, n, d; };
template <int VPL>
__global__ void __launch_bounds__(256) perceiver_prep_kernel(PrepArgs a) {
  pdl_enter();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= a.B * a.T * a.n) return;
  constexpr int d = VPL * 128;
  const int t = (warp / a.n) % a.T, b = warp / (a.n * a.T);
  const float mk = a.mask ? a.mask[b * a.T + t] : 1.f;
  float4 v[VPL];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 x = *reinterpret_cast<const float4*>(a.x + (size_t)warp * d + c);
    const float4 e = *reinterpret_cast<const float4*>(a.tpe + (size_t)t * d + c);
    v[i] = make_float4(x.x + e.x * mk, x.y + e.y * mk, x.z + e.z * mk, x.w + e.w * mk);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
    q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
#pragma unroll
  for (int i = 0; i < VPL; ++i)
    *reinterpret_cast<float4*>(a.xhat + (size_t)warp * d + (i * 32 + lane) * 4) =
        make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
}

// L[b*Q + q, :] = latents[q, :]
__global__ void perceiver_init_kernel(const float* __restrict__ latents, float* __restrict__ L, int B, int Q, int d) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * Q * d) L[idx] = latents[idx % (Q * d)];
}

// commit time: value-side operand with the norm_media weight folded in (split bf16): wvx[(h*dh + c), m] = Wv[h*dh + c, m] * g[m]
__global__ void perceiver_fold_kernel(const float* __restrict__ Wv, const float* __restrict__ g, __nv_bfloat16* __restrict__ wvx, int d, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * 64 * d) return;
  const int m = idx % d, c = (idx / d) % 64, h = idx / (64 * d);
  __nv_bfloat16 hi, lo;
  split_bf16(Wv[(size_t)(h * 64 + c) * d + m] * g[m], hi, lo);
  wvx[(size_t)(h * 64 + c) * 2 * d + m] = hi; wvx[(size_t)(h * 64 + c) * 2 * d + d + m] = lo;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Score path in exact fp32 (the softmax can be saturated: score errors are amplified, value errors are not).  CTA = (16 latent
// rows, head h):   q_h = scale * lhat Wq_h^T,  k_h = lhat Wk_h^T   (written to the qkv buffer for the latent-key scores),
//                  qtilde_h = g (.) Wk_h^T q_h  (the query in feature space),  cq = q_h . (Wk_h b)
struct QPathArgs {
  const float* lhat;                    // (Mq, d) LayerNorm(latents)
  const float *Wq, *Wk, *g, *kb;        // (inner, d) x2, norm_media weight (d), Wk b (inner)
  float* qkv; int ldq; int inner;       // q -> cols [h*64, +64), k -> cols [inner + h*64, +64)
  float* qt; float* cq;                 // qt rows (h*Mp + row), d floats; cq (Mq, H)
  int Mq, Mp, H, d; float scale;
};
constexpr int QP_ROWS = 16, QP_MC = 16;   // static shared memory stays under 48 KB
template <int D>
__global__ void __launch_bounds__(256) perceiver_qpath_kernel(QPathArgs a) {
  __shared__ __align__(16) float sl[QP_ROWS][D + 4];
  __shared__ __align__(16) float sw[2][128][QP_MC + 4];    // double-buffered weight chunk: rows 0..63 Wq_h, 64..127 Wk_h (cp.async)
  float (*sqk)[128] = reinterpret_cast<float (*)[128]>(&sw[0][0][0]);     // q_h (scaled) | k_h: reuses the chunk buffers after step 1
  static_assert(sizeof(float) * QP_ROWS * 128 <= sizeof(sw), "sqk must fit into the chunk buffers");
  pdl_enter();
  const int r0 = blockIdx.x * QP_ROWS, h = blockIdx.y, tid = threadIdx.x;
  auto load_chunk = [&](int buf, int m0) {
    for (int e = tid; e < 128 * (QP_MC / 4); e += 256) {
      const int wr = e / (QP_MC / 4), mc = (e % (QP_MC / 4)) * 4;
      const float* W = wr < 64 ? a.Wq : a.Wk;
      cp_async16(&sw[buf][wr][mc], W + (size_t)(h * 64 + (wr & 63)) * D + m0 + mc);
    }
  };
  load_chunk(0, 0);
  cp_async_commit();
  for (int e = tid; e < QP_ROWS * (D / 4); e += 256) {
    const int r = e / (D / 4), c = (e % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < a.Mq) v = *reinterpret_cast<const float4*>(a.lhat + (size_t)(r0 + r) * D + c);
    *reinterpret_cast<float4*>(&sl[r][c]) = v;
  }
  // step 1: thread = (output column c of [q | k], row group of 8); the next weight chunk streams in while this one is consumed
  const int c = tid & 127, rg = tid >> 7;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  int buf = 0;
  for (int m0 = 0; m0 < D; m0 += QP_MC, buf ^= 1) {
    if (m0 + QP_MC < D) load_chunk(buf ^ 1, m0 + QP_MC);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
#pragma unroll
    for (int m = 0; m < QP_MC; m += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&sw[buf][c][m]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 l = *reinterpret_cast<const float4*>(&sl[rg * 8 + k][m0 + m]);
        acc[k] = fmaf(l.x, w.x, acc[k]); acc[k] = fmaf(l.y, w.y, acc[k]); acc[k] = fmaf(l.z, w.z, acc[k]); acc[k] = fmaf(l.w, w.w, acc[k]);
      }
    }
    __syncthreads();                 // everyone is done with `buf` before the iteration after next refills it
  }
  cp_async_wait<0>();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = rg * 8 + k;
    const float v = c < 64 ? acc[k] * a.scale : acc[k];
    sqk[r][c] = v;
    if (r0 + r < a.Mq) a.qkv[(size_t)(r0 + r) * a.ldq + (c < 64 ? 0 : a.inner) + h * 64 + (c & 63)] = v;
  }
  __syncthreads();
  // cq[row, h] = q_h . kb_h
  if (tid < QP_ROWS && r0 + tid < a.Mq) {
    float s = 0.f;
    for (int j = 0; j < 64; ++j) s = fmaf(sqk[tid][j], a.kb[h * 64 + j], s);
    a.cq[(size_t)(r0 + tid) * a.H + h] = s;
  }
  // step 2: qtilde[r, m] = g[m] * sum_j q_h[r, j] Wk[h*64 + j, m]      thread = feature-space column m (coalesced Wk reads)
  for (int m = tid; m < D; m += 256) {
    float o[QP_ROWS];
#pragma unroll
    for (int r = 0; r < QP_ROWS; ++r) o[r] = 0.f;
    for (int j = 0; j < 64; ++j) {
      const float w = a.Wk[(size_t)(h * 64 + j) * D + m];
#pragma unroll
      for (int r = 0; r < QP_ROWS; ++r) o[r] = fmaf(sqk[r][j], w, o[r]);
    }
    const float gm = a.g[m];
#pragma unroll
    for (int r = 0; r < QP_ROWS; ++r)
      if (r0 + r < a.Mq) a.qt[((size_t)h * a.Mp + r0 + r) * D + m] = o[r] * gm;
  }
}

// ---- warp-level tensor-core helpers for the two passes over xhat (3xTF32: fp32 operands split into tf32 hi + lo, three
// mma.sync.m16n8k8 per product with fp32 accumulation -> ~2^-21 relative operand error, i.e. fp32-grade scores) ----
__device__ __forceinline__ void tf32_split(float v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(v - __uint_as_float(hi)));
}
// D += A(16x8, row) . B(8x8, col); fragment layout (g = lane / 4, t = lane % 4): a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// b0 (k = t, n = g) b1 (k = t+4, n = g); d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_3x(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_tf32(d, al, bh0, bh1);      // small terms first
  mma_tf32(d, ah, bl0, bl1);
  mma_tf32(d, ah, bh0, bh1);
}

// ---- q-path on the warp-level tensor cores (3xTF32, fp32-grade like the SIMT kernel above, which stays as MDTB200_PERC_QPATH=simt) ----
// (1) [q | k] = lhat . [Wq ; Wk]^T : CTA = 64 latent rows x one 64-column block (columns of q for blockIdx.y < H, of k otherwise),
//     warp = 16 rows x 64 columns (8 n-tiles); A and B fragments are streamed from global memory as 16-byte loads with the k-permutation
//     of the scores kernel (lane t owns reduction indices 4t..4t+3 of a 16-wide chunk = k (t, t+4) of two consecutive k-steps).
//     q is scaled and its score constant cq[row, h] = q_h . (Wk_h b) is reduced from the accumulators.
__global__ void __launch_bounds__(128) perceiver_qk_mma_kernel(QPathArgs a) {
  pdl_enter();
  const int d = a.d, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int cb = blockIdx.y, is_k = cb >= a.H, h = is_k ? cb - a.H : cb;
  const int r_a = blockIdx.x * 64 + warp * 16 + g, r_b = r_a + 8;
  const float* W = (is_k ? a.Wk : a.Wq) + (size_t)h * 64 * d;
  const float* xa = a.lhat + (size_t)(r_a < a.Mq ? r_a : a.Mq - 1) * d + 4 * t;
  const float* xb = a.lhat + (size_t)(r_b < a.Mq ? r_b : a.Mq - 1) * d + 4 * t;
  const float* wp = W + (size_t)g * d + 4 * t;
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll 2
  for (int j = 0; j < d / 16; ++j) {
    const float4 va = __ldg(reinterpret_cast<const float4*>(xa + 16 * j)), vb = __ldg(reinterpret_cast<const float4*>(xb + 16 * j));
    float4 wv[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) wv[n] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)n * 8 * d + 16 * j));
    uint32_t ah0[4], al0[4], ah1[4], al1[4];
    tf32_split(va.x, ah0[0], al0[0]); tf32_split(vb.x, ah0[1], al0[1]); tf32_split(va.y, ah0[2], al0[2]); tf32_split(vb.y, ah0[3], al0[3]);
    tf32_split(va.z, ah1[0], al1[0]); tf32_split(vb.z, ah1[1], al1[1]); tf32_split(va.w, ah1[2], al1[2]); tf32_split(vb.w, ah1[3], al1[3]);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      uint32_t bh[4], bl[4];
      tf32_split(wv[n].x, bh[0], bl[0]); tf32_split(wv[n].y, bh[1], bl[1]); tf32_split(wv[n].z, bh[2], bl[2]); tf32_split(wv[n].w, bh[3], bl[3]);
      mma_3x(acc[n], ah0, al0, bh[0], bh[1], bl[0], bl[1]);
      mma_3x(acc[n], ah1, al1, bh[2], bh[3], bl[2], bl[3]);
    }
  }
  float cqa = 0.f, cqb = 0.f;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = n * 8 + 2 * t + (i & 1), row = (i & 2) ? r_b : r_a;
      const float v = is_k ? acc[n][i] : acc[n][i] * a.scale;
      if (!is_k) { const float kbv = a.kb[h * 64 + c]; if (i & 2) cqb = fmaf(v, kbv, cqb); else cqa = fmaf(v, kbv, cqa); }
      if (row < a.Mq) a.qkv[(size_t)row * a.ldq + (is_k ? a.inner : 0) + h * 64 + c] = v;
    }
  }
  if (!is_k) {
    cqa += __shfl_xor_sync(0xffffffffu, cqa, 1); cqa += __shfl_xor_sync(0xffffffffu, cqa, 2);
    cqb += __shfl_xor_sync(0xffffffffu, cqb, 1); cqb += __shfl_xor_sync(0xffffffffu, cqb, 2);
    if (t == 0) {
      if (r_a < a.Mq) a.cq[(size_t)r_a * a.H + h] = cqa;
      if (r_b < a.Mq) a.cq[(size_t)r_b * a.H + h] = cqb;
    }
  }
}
// (2) qtilde_h = g (.) (q_h . Wk_h) : CTA = 64 rows x 128 feature-space columns of head h (grid (Mq/64, d/128, H)), warp = 16 rows x 128
//     columns = 4 groups of four n-tiles whose B fragments come from one 16-byte load per k row (lane g owns columns 4g..4g+3 of a
//     32-column group: column 4g + u is index g of n-tile u).  K = 64 (the head dimension).
__global__ void __launch_bounds__(128) perceiver_qt_mma_kernel(QPathArgs a) {
  pdl_enter();
  const int d = a.d, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.z, c0 = blockIdx.y * 128;
  const int r_a = blockIdx.x * 64 + warp * 16 + g, r_b = r_a + 8;
  const float* qa = a.qkv + (size_t)(r_a < a.Mq ? r_a : a.Mq - 1) * a.ldq + h * 64 + 4 * t;
  const float* qb = a.qkv + (size_t)(r_b < a.Mq ? r_b : a.Mq - 1) * a.ldq + h * 64 + 4 * t;
  const float* wk = a.Wk + (size_t)(h * 64 + 4 * t) * d + c0 + 4 * g;          // k row 4t (+0..3) of the chunk, columns 4g..4g+3 of group 0
  float acc[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {                 // 16 reduction indices per chunk
    const float4 va = *reinterpret_cast<const float4*>(qa + 16 * j), vb = *reinterpret_cast<const float4*>(qb + 16 * j);
    uint32_t ah0[4], al0[4], ah1[4], al1[4];
    tf32_split(va.x, ah0[0], al0[0]); tf32_split(vb.x, ah0[1], al0[1]); tf32_split(va.y, ah0[2], al0[2]); tf32_split(vb.y, ah0[3], al0[3]);
    tf32_split(va.z, ah1[0], al1[0]); tf32_split(vb.z, ah1[1], al1[1]); tf32_split(va.w, ah1[2], al1[2]); tf32_split(vb.w, ah1[3], al1[3]);
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      float4 w[4];                               // k rows 4t + 0..3 of this chunk
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) w[kk] = __ldg(reinterpret_cast<const float4*>(wk + (size_t)(16 * j + kk) * d + grp * 32));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float b00 = u == 0 ? w[0].x : u == 1 ? w[0].y : u == 2 ? w[0].z : w[0].w;      // k = t   (row 4t)     step 0
        const float b01 = u == 0 ? w[1].x : u == 1 ? w[1].y : u == 2 ? w[1].z : w[1].w;      // k = t+4 (row 4t + 1) step 0
        const float b10 = u == 0 ? w[2].x : u == 1 ? w[2].y : u == 2 ? w[2].z : w[2].w;      // step 1
        const float b11 = u == 0 ? w[3].x : u == 1 ? w[3].y : u == 2 ? w[3].z : w[3].w;
        uint32_t h0, l0, h1, l1;
        tf32_split(b00, h0, l0); tf32_split(b01, h1, l1);
        mma_3x(acc[grp * 4 + u], ah0, al0, h0, h1, l0, l1);
        tf32_split(b10, h0, l0); tf32_split(b11, h1, l1);
        mma_3x(acc[grp * 4 + u], ah1, al1, h0, h1, l0, l1);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < 16; ++n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = c0 + (n >> 2) * 32 + 4 * (2 * t + (i & 1)) + (n & 3), row = (i & 2) ? r_b : r_a;
      if (row < a.Mq) a.qt[((size_t)h * a.Mp + row) * d + col] = acc[n][i] * a.g[col];
    }
  }
}

// scores[b, r, f] = qtilde[r] . xhat[b, f] + cq[r]      r = h*Q + q < HQ <= 8 NT.   CTA = 64 features of one sample, warp = 16 of
// them (one m-tile); the feature-space queries sit in shared memory as tf32 hi / lo planes, the features stream from global memory
// straight into the A fragments (each xhat element is used by exactly one warp) as 16-byte loads: within a 16-column chunk lane t
// owns columns 4t..4t+3, which are the k indices (t, t+4) of two consecutive k-steps -- the same permutation is applied to B.
struct ScoreArgs { const float* qt; const float* cq; const float* xhat; float* scores; int B, F, Fp, Q, H, Mp, d; };
constexpr int SC_FPC = 128, SC_THREADS = 256;       // 8 warps share one copy of the query planes (2 CTAs = 16 warps per SM)
inline int attn_nt(int HQ) { const int n = (HQ + 7) / 8; return n <= 4 ? n : n <= 6 ? 6 : 8; }     // instantiated n-tile counts: 1 2 3 4 6 8
inline int score_dp(int d) { return d + 16; }                         // row stride = 16 (mod 32) words: conflict-free LDS.128 fragments
inline size_t score_smem_bytes(int HQ, int d) { return (size_t)2 * attn_nt(HQ) * 8 * score_dp(d) * sizeof(float); }
template <int NT>
__global__ void __launch_bounds__(SC_THREADS) perceiver_scores_kernel(ScoreArgs a) {
  extern __shared__ __align__(16) float sc_smem[];
  pdl_enter();
  const int d = a.d, DP = d + 16, D4 = d / 4, HQ = a.H * a.Q, NR = NT * 8;
  uint32_t* qh = reinterpret_cast<uint32_t*>(sc_smem);      // [NR][DP] tf32 hi
  uint32_t* ql = qh + NR * DP;                             // [NR][DP] tf32 lo
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  for (int e0 = 0; e0 < NR * D4; e0 += SC_THREADS * 6) {         // batches of 6 independent 16-byte loads per thread
    float4 v[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int e = e0 + u * SC_THREADS + tid, r = e / D4, c = (e % D4) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e < NR * D4 && r < HQ) v[u] = *reinterpret_cast<const float4*>(a.qt + ((size_t)(r / a.Q) * a.Mp + (size_t)b * a.Q + r % a.Q) * d + c);
    }
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int e = e0 + u * SC_THREADS + tid, r = e / D4, c = (e % D4) * 4;
      if (e < NR * D4) {
        uint4 hi, lo;
        tf32_split(v[u].x, hi.x, lo.x); tf32_split(v[u].y, hi.y, lo.y); tf32_split(v[u].z, hi.z, lo.z); tf32_split(v[u].w, hi.w, lo.w);
        *reinterpret_cast<uint4*>(qh + r * DP + c) = hi;
        *reinterpret_cast<uint4*>(ql + r * DP + c) = lo;
      }
    }
  }
  __syncthreads();
  const int f_a = blockIdx.x * SC_FPC + warp * 16 + g, f_b = f_a + 8;
  if (blockIdx.x * SC_FPC + warp * 16 >= a.F) return;
  const float* xa = a.xhat + ((size_t)b * a.F + (f_a < a.F ? f_a : a.F - 1)) * d + 4 * t;
  const float* xb = a.xhat + ((size_t)b * a.F + (f_b < a.F ? f_b : a.F - 1)) * d + 4 * t;
  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n) acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f;
  const uint32_t* bh = qh + g * DP + 4 * t;
  const uint32_t* bl = ql + g * DP + 4 * t;
  // register ring of SC_PF chunks: the loads of chunk j + SC_PF are issued as soon as chunk j has been consumed, so SC_PF
  // 16-byte loads per row stay in flight per thread (the loop is otherwise bound by one global-memory latency per chunk)
  constexpr int SC_PF = 6;
  const int nj = d / 16;
  float4 ra[SC_PF], rb[SC_PF];
#pragma unroll
  for (int u = 0; u < SC_PF; ++u) {
    ra[u] = u < nj ? __ldg(reinterpret_cast<const float4*>(xa + 16 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
    rb[u] = u < nj ? __ldg(reinterpret_cast<const float4*>(xb + 16 * u)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j0 = 0; j0 < nj; j0 += SC_PF) {
#pragma unroll
    for (int u = 0; u < SC_PF; ++u) {
      const int j = j0 + u;
      if (j >= nj) break;
      const float4 va = ra[u], vb = rb[u];
      if (j + SC_PF < nj) {
        ra[u] = __ldg(reinterpret_cast<const float4*>(xa + 16 * (j + SC_PF)));
        rb[u] = __ldg(reinterpret_cast<const float4*>(xb + 16 * (j + SC_PF)));
      }
      uint32_t ah0[4], al0[4], ah1[4], al1[4];
      tf32_split(va.x, ah0[0], al0[0]); tf32_split(vb.x, ah0[1], al0[1]); tf32_split(va.y, ah0[2], al0[2]); tf32_split(vb.y, ah0[3], al0[3]);
      tf32_split(va.z, ah1[0], al1[0]); tf32_split(vb.z, ah1[1], al1[1]); tf32_split(va.w, ah1[2], al1[2]); tf32_split(vb.w, ah1[3], al1[3]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint4 h4 = *reinterpret_cast<const uint4*>(bh + n * 8 * DP + 16 * j);
        const uint4 l4 = *reinterpret_cast<const uint4*>(bl + n * 8 * DP + 16 * j);
        mma_3x(acc[n], ah0, al0, h4.x, h4.y, l4.x, l4.y);
        mma_3x(acc[n], ah1, al1, h4.z, h4.w, l4.z, l4.w);
      }
    }
  }
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = n * 8 + 2 * t + (i & 1), f = (i & 2) ? f_b : f_a;
      if (r < HQ && f < a.F) a.scores[((size_t)b * HQ + r) * a.Fp + f] = acc[n][i] + a.cq[((size_t)b * a.Q + r % a.Q) * a.H + r / a.Q];
    }
  }
}

// softmax over the F feature keys + Q latent keys (perceiver_resampler.py:69-79), then
//   z[r, cols] = sum_f alpha[r, f] xhat[f, cols]  -> split-bf16 operand (head-major rows), wsum[row, h] = sum_f alpha[r, f],
//   olat[row, h*64 + c] = sum_j alpha[r, F + j] v_lat[j, h*64 + c]
// CTA = (128-column tile, sample), warp = 32 columns = two m-tiles of z^T[c, r] = sum_f xhat[f, c] alpha[r, f]: lane (g, t) loads
// xhat[k0 + t (+4)][c_w + 4g .. +3] (full 128-byte lines per warp); its four columns are rows (g, g+8) of the two m-tiles.
struct ZArgs {
  const float* scores; const float* qkv; int ldq; int inner; const float* xhat;
  __nv_bfloat16* z16; float* wsum; float* olat; int B, F, Fp, Q, H, Mp, d;
};
constexpr int Z_CT = 256, Z_THREADS = 256;
inline int z_sp(int F, int Q) { return (F + Q + 3) / 8 * 8 + 4; }     // row stride = 4 (mod 8) words >= F + Q: conflict-free LDS.32 B fragments
inline size_t z_smem_bytes(int HQ, int F, int Q) { return (size_t)2 * attn_nt(HQ) * 8 * z_sp(F, Q) * sizeof(float); }
template <int NT>
__global__ void __launch_bounds__(Z_THREADS, 2) perceiver_softmax_z_kernel(ZArgs a) {
  extern __shared__ __align__(16) float z_smem[];
  pdl_enter();
  const int HQ = a.H * a.Q, NK = a.F + a.Q, SP = (NK + 3) / 8 * 8 + 4, d = a.d, NR = NT * 8;
  float* sa = z_smem;                                        // [NR][SP] scores -> probabilities -> tf32 hi
  uint32_t* sl = reinterpret_cast<uint32_t*>(sa + NR * SP);  // [NR][SP] tf32 lo
  const int b = blockIdx.y, c0 = blockIdx.x * Z_CT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int F4 = a.F / 4;
  for (int e = tid; e < HQ * F4; e += Z_THREADS) cp_async16(sa + (e / F4) * SP + (e % F4) * 4, a.scores + ((size_t)b * HQ + e / F4) * a.Fp + (e % F4) * 4);
  for (int e = tid; e < HQ * (a.F - F4 * 4); e += Z_THREADS) {
    const int r = e / (a.F - F4 * 4), f = F4 * 4 + e % (a.F - F4 * 4);
    sa[r * SP + f] = a.scores[((size_t)b * HQ + r) * a.Fp + f];
  }
  for (int e = tid; e < HQ * a.Q; e += Z_THREADS) {          // latent keys: q_hq . k_lat[j, h]  (the scale is already inside q)
    const int r = e / a.Q, j = e % a.Q, h = r / a.Q, qi = r % a.Q;
    const float4* qp = reinterpret_cast<const float4*>(a.qkv + ((size_t)b * a.Q + qi) * a.ldq + h * 64);
    const float4* kp = reinterpret_cast<const float4*>(a.qkv + ((size_t)b * a.Q + j) * a.ldq + a.inner + h * 64);
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { const float4 x = qp[c], y = kp[c]; acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc); }
    sa[r * SP + a.F + j] = acc;
  }
  cp_async_wait_all();
  __syncthreads();
  for (int r = warp; r < NR; r += Z_THREADS / 32) {          // warp per row: max, exp, sum; then the tf32 planes (zero padding rows / keys)
    float* row = sa + r * SP;
    uint32_t* rlo = sl + r * SP;
    if (r >= HQ) { for (int k = lane; k < SP; k += 32) { row[k] = 0.f; rlo[k] = 0u; } continue; }
    float mx = -INFINITY;
    for (int k = lane; k < NK; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < NK; k += 32) { const float ex = expf(row[k] - mx); row[k] = ex; sum += ex; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float ws = 0.f;
    for (int k = lane; k < SP; k += 32) {
      const float p = k < NK ? row[k] * inv : 0.f;
      if (k < a.F) {
        ws += p;
        uint32_t hi, lo;
        tf32_split(p, hi, lo);
        row[k] = __uint_as_float(hi); rlo[k] = lo;
      } else {
        row[k] = p; rlo[k] = 0u;           // latent-key probabilities stay fp32 (used by the olat sum below, never by the MMAs)
      }
    }
    ws = warp_sum(ws);
    if (lane == 0 && blockIdx.x == 0) a.wsum[((size_t)b * a.Q + r % a.Q) * a.H + r / a.Q] = ws;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < HQ * 64; e += Z_THREADS) {
      const int r = e / 64, c = e % 64, h = r / a.Q, qi = r % a.Q;
      float acc = 0.f;
      for (int j = 0; j < a.Q; ++j) acc = fmaf(sa[r * SP + a.F + j], a.qkv[((size_t)b * a.Q + j) * a.ldq + 2 * a.inner + h * 64 + c], acc);
      a.olat[((size_t)b * a.Q + qi) * a.inner + h * 64 + c] = acc;
    }
  }
  const int cw = c0 + warp * 32;
  if (cw >= d) return;
  float acc[2][NT][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
  const float* xp = a.xhat + (size_t)b * a.F * d + cw + 4 * g;
  const uint32_t* bh = reinterpret_cast<const uint32_t*>(sa) + g * SP + t;
  const uint32_t* bl = sl + g * SP + t;
  const int nks = (a.F + 7) / 8;
  constexpr int Z_PF = 5;                                    // register ring, as in the scores kernel
  auto xrow = [&](int k) { return reinterpret_cast<const float4*>(xp + (size_t)(k < a.F ? k : a.F - 1) * d); };
  float4 r1[Z_PF], r2[Z_PF];
#pragma unroll
  for (int u = 0; u < Z_PF; ++u) { r1[u] = __ldg(xrow(u * 8 + t)); r2[u] = __ldg(xrow(u * 8 + t + 4)); }
  for (int ks0 = 0; ks0 < nks; ks0 += Z_PF) {
#pragma unroll
    for (int u = 0; u < Z_PF; ++u) {
      const int ks = ks0 + u;
      if (ks >= nks) break;
      const int k1 = ks * 8 + t, k2 = k1 + 4;               // keys past F contribute nothing (alpha forced to 0 below)
      const float4 v1 = r1[u], v2 = r2[u];
      if (ks + Z_PF < nks) { r1[u] = __ldg(xrow(k1 + 8 * Z_PF)); r2[u] = __ldg(xrow(k2 + 8 * Z_PF)); }
      uint32_t ah0[4], al0[4], ah1[4], al1[4];
      tf32_split(v1.x, ah0[0], al0[0]); tf32_split(v1.y, ah0[1], al0[1]); tf32_split(v2.x, ah0[2], al0[2]); tf32_split(v2.y, ah0[3], al0[3]);
      tf32_split(v1.z, ah1[0], al1[0]); tf32_split(v1.w, ah1[1], al1[1]); tf32_split(v2.z, ah1[2], al1[2]); tf32_split(v2.w, ah1[3], al1[3]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t h0 = k1 < a.F ? bh[n * 8 * SP + ks * 8] : 0u, h1 = k2 < a.F ? bh[n * 8 * SP + ks * 8 + 4] : 0u;
        const uint32_t l0 = k1 < a.F ? bl[n * 8 * SP + ks * 8] : 0u, l1 = k2 < a.F ? bl[n * 8 * SP + ks * 8 + 4] : 0u;
        mma_3x(acc[0][n], ah0, al0, h0, h1, l0, l1);
        mma_3x(acc[1][n], ah1, al1, h0, h1, l0, l1);
      }
    }
  }
  // accumulator (m-tile m, n-tile n): rows g / g+8 are columns cw + 4g + 2m (+1), columns 2t (+1) are the query rows r
#pragma unroll
  for (int n = 0; n < NT; ++n) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = n * 8 + 2 * t + i;
      if (r >= HQ) continue;
      const size_t row = (size_t)(r / a.Q) * a.Mp + (size_t)b * a.Q + r % a.Q;
      const float o[4] = {acc[0][n][i], acc[0][n][2 + i], acc[1][n][i], acc[1][n][2 + i]};
      __align__(8) __nv_bfloat16 hi[4];
      __align__(8) __nv_bfloat16 lo[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) split_bf16(o[u], hi[u], lo[u]);
      __nv_bfloat16* po = a.z16 + row * 2 * d + cw + 4 * g;
      *reinterpret_cast<uint2*>(po) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(po + d) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

template <int NT>
inline cudaError_t attn_configure_nt(size_t ss, size_t zs) {
  cudaError_t e = cudaFuncSetAttribute(perceiver_scores_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(perceiver_softmax_z_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zs);
}
inline cudaError_t attn_configure(int HQ, size_t ss, size_t zs) {
  switch (attn_nt(HQ)) {
    case 1: return attn_configure_nt<1>(ss, zs);
    case 2: return attn_configure_nt<2>(ss, zs);
    case 3: return attn_configure_nt<3>(ss, zs);
    case 4: return attn_configure_nt<4>(ss, zs);
    case 6: return attn_configure_nt<6>(ss, zs);
    default: return attn_configure_nt<8>(ss, zs);
  }
}
template <int NT>
inline void attn_launch_nt(const ScoreArgs& sa, const ZArgs& za, cudaStream_t st) {
  const int HQ = sa.H * sa.Q;
  launch_pdl(perceiver_scores_kernel<NT>, dim3((sa.F + SC_FPC - 1) / SC_FPC, sa.B), dim3(SC_THREADS), score_smem_bytes(HQ, sa.d), st, sa);
  launch_pdl(perceiver_softmax_z_kernel<NT>, dim3((sa.d + Z_CT - 1) / Z_CT, sa.B), dim3(Z_THREADS), z_smem_bytes(HQ, sa.F, sa.Q), st, za);
}
inline void attn_launch(const ScoreArgs& sa, const ZArgs& za, cudaStream_t st) {
  switch (attn_nt(sa.H * sa.Q)) {
    case 1: attn_launch_nt<1>(sa, za, st); break;
    case 2: attn_launch_nt<2>(sa, za, st); break;
    case 3: attn_launch_nt<3>(sa, za, st); break;
    case 4: attn_launch_nt<4>(sa, za, st); break;
    case 6: attn_launch_nt<6>(sa, za, st); break;
    default: attn_launch_nt<8>(sa, za, st); break;
  }
}

// o[row, h*64 + c] = of[(h*Mp + row), c] + wsum[row, h] * vb[h*64 + c] + olat[row, h*64 + c]  -> split-bf16 operand of to_out
struct CombineArgs { const float* of; const float* wsum; const float* vb; const float* olat; __nv_bfloat16* o16; int Mq, Mp, H, inner; };
__global__ void __launch_bounds__(256) perceiver_combine_kernel(CombineArgs a) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.Mq * a.inner) return;
  const int c = idx % a.inner, row = idx / a.inner, h = c / 64;
  const float v = a.of[((size_t)h * a.Mp + row) * 64 + c % 64] + a.wsum[(size_t)row * a.H + h] * a.vb[c] + a.olat[idx];
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  a.o16[(size_t)row * 2 * a.inner + c] = hi;
  a.o16[(size_t)row * 2 * a.inner + a.inner + c] = lo;
}

}}  // namespace mdt::pr
