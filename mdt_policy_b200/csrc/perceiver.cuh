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
  __shared__ __align__(16) float sw[128][QP_MC + 4];       // rows 0..63: Wq_h chunk, 64..127: Wk_h chunk
  __shared__ float sqk[QP_ROWS][128];                      // q_h (scaled) | k_h
  pdl_enter();
  const int r0 = blockIdx.x * QP_ROWS, h = blockIdx.y, tid = threadIdx.x;
  for (int e = tid; e < QP_ROWS * (D / 4); e += 256) {
    const int r = e / (D / 4), c = (e % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < a.Mq) v = *reinterpret_cast<const float4*>(a.lhat + (size_t)(r0 + r) * D + c);
    *reinterpret_cast<float4*>(&sl[r][c]) = v;
  }
  // step 1: thread = (output column c of [q | k], row group of 8)
  const int c = tid & 127, rg = tid >> 7;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int m0 = 0; m0 < D; m0 += QP_MC) {
    __syncthreads();
    for (int e = tid; e < 128 * (QP_MC / 4); e += 256) {
      const int wr = e / (QP_MC / 4), mc = (e % (QP_MC / 4)) * 4;
      const float* W = wr < 64 ? a.Wq : a.Wk;
      *reinterpret_cast<float4*>(&sw[wr][mc]) = *reinterpret_cast<const float4*>(W + (size_t)(h * 64 + (wr & 63)) * D + m0 + mc);
    }
    __syncthreads();
#pragma unroll 4
    for (int m = 0; m < QP_MC; m += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&sw[c][m]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 l = *reinterpret_cast<const float4*>(&sl[rg * 8 + k][m0 + m]);
        acc[k] = fmaf(l.x, w.x, acc[k]); acc[k] = fmaf(l.y, w.y, acc[k]); acc[k] = fmaf(l.z, w.z, acc[k]); acc[k] = fmaf(l.w, w.w, acc[k]);
      }
    }
  }
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

// asynchronous 16-byte global -> shared copies (LDGSTS): every copy of a tile is in flight at once, no register staging
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// scores[b, r, f] = qtilde[r] . xhat[b, f] + cq[r]      r = h*Q + q < HQ <= 64; CTA = (SC_CH features of one sample): the HQ feature-space
// queries are loaded once, the features stream through two 32-row tiles filled by cp.async while the previous tile is consumed.
struct ScoreArgs { const float* qt; const float* cq; const float* xhat; float* scores; int B, F, Fp, Q, H, Mp, d; };
constexpr int SC_FT = 32, SC_CH = 128;
inline size_t score_smem_bytes(int HQ, int d) { return (size_t)(HQ + 2 * SC_FT) * (d + 4) * sizeof(float); }
__global__ void __launch_bounds__(256) perceiver_scores_kernel(ScoreArgs a) {
  extern __shared__ __align__(16) float sc_smem[];
  pdl_enter();
  const int d = a.d, DP = d + 4, D4 = d / 4, HQ = a.H * a.Q;
  float* sq = sc_smem;                 // [HQ][DP]
  float* sx = sq + HQ * DP;            // [2][SC_FT][DP]
  const int b = blockIdx.y, fbeg = blockIdx.x * SC_CH, tid = threadIdx.x;
  const int fend = fbeg + SC_CH < a.F ? fbeg + SC_CH : a.F;
  auto load_tile = [&](int buf, int f0) {
    float* dst = sx + buf * SC_FT * DP;
    for (int e = tid; e < SC_FT * D4; e += 256) {
      const int fl = e / D4, c = (e % D4) * 4;
      if (f0 + fl < a.F) cp_async16(dst + fl * DP + c, a.xhat + ((size_t)b * a.F + f0 + fl) * d + c);
      else *reinterpret_cast<float4*>(dst + fl * DP + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  for (int e = tid; e < HQ * D4; e += 256) {
    const int r = e / D4, c = (e % D4) * 4, h = r / a.Q, qi = r % a.Q;
    cp_async16(sq + r * DP + c, a.qt + ((size_t)h * a.Mp + (size_t)b * a.Q + qi) * d + c);
  }
  load_tile(0, fbeg);
  cp_async_commit();
  const int fl = tid & 31, rg = tid >> 5;          // lanes = features (conflict-free rows), warp = row group (broadcast q rows)
  int buf = 0;
  for (int f0 = fbeg; f0 < fend; f0 += SC_FT, buf ^= 1) {
    if (f0 + SC_FT < fend) load_tile(buf ^ 1, f0 + SC_FT);
    cp_async_commit();
    cp_async_wait<1>();                 // the tile requested one iteration ago (and the queries) have landed
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    const float* xp = sx + buf * SC_FT * DP + fl * DP;
    for (int c = 0; c < d; c += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(xp + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = rg + 8 * k;
        if (r < HQ) {
          const float4 qv = *reinterpret_cast<const float4*>(sq + r * DP + c);
          acc[k] = fmaf(xv.x, qv.x, acc[k]); acc[k] = fmaf(xv.y, qv.y, acc[k]); acc[k] = fmaf(xv.z, qv.z, acc[k]); acc[k] = fmaf(xv.w, qv.w, acc[k]);
        }
      }
    }
    if (f0 + fl < a.F) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = rg + 8 * k;
        if (r < HQ) a.scores[((size_t)b * HQ + r) * a.Fp + f0 + fl] = acc[k] + a.cq[((size_t)b * a.Q + r % a.Q) * a.H + r / a.Q];
      }
    }
    __syncthreads();                    // everyone is done with `buf` before the next iteration refills it
  }
  cp_async_wait<0>();
}

// softmax over the F feature keys + Q latent keys (perceiver_resampler.py:69-79), then
//   z[r, cols] = sum_f alpha[r, f] xhat[f, cols]  -> split-bf16 operand (head-major rows), wsum[row, h] = sum_f alpha[r, f],
//   olat[row, h*64 + c] = sum_j alpha[r, F + j] v_lat[j, h*64 + c]                      CTA = (96-column tile, sample)
struct ZArgs {
  const float* scores; const float* qkv; int ldq; int inner; const float* xhat;
  __nv_bfloat16* z16; float* wsum; float* olat; int B, F, Fp, Q, H, Mp, d;
};
constexpr int Z_CT = 96, Z_FT = 56;       // 392 = 7 x 56 feature rows per tile
inline size_t z_smem_bytes(int HQ, int F, int Q) { return ((size_t)HQ * (F + Q + 1) + (size_t)2 * Z_FT * (Z_CT + 4)) * sizeof(float); }
__global__ void __launch_bounds__(256) perceiver_softmax_z_kernel(ZArgs a) {
  extern __shared__ __align__(16) float z_smem[];
  pdl_enter();
  const int HQ = a.H * a.Q, NK = a.F + a.Q, SP = NK + 1, d = a.d;
  float* sa = z_smem;                       // [HQ][SP] scores -> probabilities
  float* sx = sa + (size_t)((HQ * SP + 3) / 4) * 4;     // [2][Z_FT][Z_CT + 4], 16-byte aligned
  const int b = blockIdx.y, c0 = blockIdx.x * Z_CT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto load_tile = [&](int buf, int f0) {
    float* dst = sx + buf * Z_FT * (Z_CT + 4);
    for (int e = tid; e < Z_FT * (Z_CT / 4); e += 256) {
      const int fl = e / (Z_CT / 4), cc = (e % (Z_CT / 4)) * 4;
      if (f0 + fl < a.F) cp_async16(dst + fl * (Z_CT + 4) + cc, a.xhat + ((size_t)b * a.F + f0 + fl) * d + c0 + cc);
      else *reinterpret_cast<float4*>(dst + fl * (Z_CT + 4) + cc) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  load_tile(0, 0);                          // the first feature tile streams in while the softmax is computed
  cp_async_commit();
  for (int e = tid; e < HQ * a.F; e += 256) sa[(e / a.F) * SP + e % a.F] = a.scores[((size_t)b * HQ + e / a.F) * a.Fp + e % a.F];
  for (int e = tid; e < HQ * a.Q; e += 256) {          // latent keys: q_hq . k_lat[j, h]  (the scale is already inside q)
    const int r = e / a.Q, j = e % a.Q, h = r / a.Q, qi = r % a.Q;
    const float* qp = a.qkv + ((size_t)b * a.Q + qi) * a.ldq + h * 64;
    const float* kp = a.qkv + ((size_t)b * a.Q + j) * a.ldq + a.inner + h * 64;
    float acc = 0.f;
    for (int c = 0; c < 64; ++c) acc = fmaf(qp[c], kp[c], acc);
    sa[r * SP + a.F + j] = acc;
  }
  __syncthreads();
  for (int r = warp; r < HQ; r += 8) {                 // warp per row: max, exp, sum
    float* row = sa + r * SP;
    float mx = -INFINITY;
    for (int k = lane; k < NK; k += 32) mx = fmaxf(mx, row[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < NK; k += 32) { const float ex = expf(row[k] - mx); row[k] = ex; sum += ex; }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    float ws = 0.f;
    for (int k = lane; k < NK; k += 32) { const float p = row[k] * inv; row[k] = p; if (k < a.F) ws += p; }
    ws = warp_sum(ws);
    if (lane == 0 && blockIdx.x == 0) a.wsum[((size_t)b * a.Q + r % a.Q) * a.H + r / a.Q] = ws;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int e = tid; e < HQ * 64; e += 256) {
      const int r = e / 64, c = e % 64, h = r / a.Q, qi = r % a.Q;
      float acc = 0.f;
      for (int j = 0; j < a.Q; ++j) acc = fmaf(sa[r * SP + a.F + j], a.qkv[((size_t)b * a.Q + j) * a.ldq + 2 * a.inner + h * 64 + c], acc);
      a.olat[((size_t)b * a.Q + qi) * a.inner + h * 64 + c] = acc;
    }
  }
  // weighted sum over the features for this column tile: thread = (float4 column, row group of 10); tiles double-buffered
  const int c4 = tid % 24, rg = tid / 24;                  // 24 float4 columns x 10 row groups (240 threads)
  float4 acc[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int buf = 0;
  for (int f0 = 0; f0 < a.F; f0 += Z_FT, buf ^= 1) {
    if (f0 + Z_FT < a.F) load_tile(buf ^ 1, f0 + Z_FT);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (rg < 10) {
      const int nf = a.F - f0 < Z_FT ? a.F - f0 : Z_FT;
      const float* tile = sx + buf * Z_FT * (Z_CT + 4);
      for (int fl = 0; fl < nf; ++fl) {
        const float4 xv = *reinterpret_cast<const float4*>(tile + fl * (Z_CT + 4) + c4 * 4);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int r = rg + 10 * k;
          if (r < HQ) {
            const float p = sa[r * SP + f0 + fl];
            acc[k].x = fmaf(p, xv.x, acc[k].x); acc[k].y = fmaf(p, xv.y, acc[k].y); acc[k].z = fmaf(p, xv.z, acc[k].z); acc[k].w = fmaf(p, xv.w, acc[k].w);
          }
        }
      }
    }
    __syncthreads();
  }
  cp_async_wait<0>();
  if (rg < 10) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const int r = rg + 10 * k;
      if (r < HQ) {
        const size_t row = (size_t)(r / a.Q) * a.Mp + (size_t)b * a.Q + r % a.Q;
        const float o[4] = {acc[k].x, acc[k].y, acc[k].z, acc[k].w};
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) split_bf16(o[t], hi[t], lo[t]);
        __nv_bfloat16* po = a.z16 + row * 2 * d + c0 + c4 * 4;
        *reinterpret_cast<uint2*>(po) = *reinterpret_cast<uint2*>(hi);
        *reinterpret_cast<uint2*>(po + d) = *reinterpret_cast<uint2*>(lo);
      }
    }
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
