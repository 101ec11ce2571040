// Exact-fp32 CUDA-core kernels of the TRAINING path (GCDenoiser.loss forward with saved activations + backward):
// generalised tiled GEMM (forward TN, dgrad NN, wgrad "reduce over rows"), column / group sums, activation forward and
// backward, LayerNorm(+AdaLN modulate) backward, tiny-sequence attention backward, gated residual forward/backward and a
// strided naive GEMM for the 7-wide action embedding / output head.  Reference math: transformer_blocks.py (Block,
// ConditionedBlock, Attention, MLP, LayerNorm), score_wrappers.py:45-63.  Deterministic: no floating-point atomics.
#pragma once
#include "kernels_simt.cuh"

namespace mdt {

// ------------------------------------------------------------------------------------------
// C[P,Q] (+)= sum_r A(p,r) * B(r,q)   128 x 64 x 16 tiles, 256 threads, 8x4 micro tiles (same inner loop as sgemm_tn_kernel).
//   A_RED_ROW = false: A is stored [P, R] (reduce dim contiguous)      true: A is stored [R, P] (reduce dim is the row)
//   B_RED_ROW = false: B is stored [Q, R]                               true: B is stored [R, Q]
//   forward   y = x W^T : A = x [M,K] (false), B = W [N,K] (false)        -> sgemm_tn_kernel (kernels_simt.cuh)
//   dgrad    dx = dy W  : A = dy [M,N] (false), B = W [N,K] as [R=N, Q=K] (true)
//   wgrad    dW = dy^T x: A = dy [M,N] as [R=M, P=N] (true), B = x [M,K] as [R=M, Q=K] (true)
struct GGemmArgs { const float* A; int lda; const float* B; int ldb; float* C; int ldc; int P, Q, R; int accumulate; };

template <bool A_RED_ROW, bool B_RED_ROW>
__global__ void __launch_bounds__(256) ggemm_kernel(GGemmArgs g) {
  __shared__ __align__(16) float As[2][16][128 + 4];
  __shared__ __align__(16) float Bs[2][16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int p0 = blockIdx.y * 128, q0 = blockIdx.x * 64;
  float4 ra0, ra1, rb;
  auto gload = [&](int r0) {
    if (A_RED_ROW) {       // tile rows = reduce index, 128 contiguous p: 16 x 128 floats = 512 float4, 2 per thread
      const int rr = tid >> 5, c = (tid & 31) * 4;
      const int ra = r0 + rr, rb2 = r0 + 8 + rr;
      auto ld = [&](int r, int pc) {
        float4 v = make_float4(0, 0, 0, 0);
        if (r < g.R) {
          const float* src = g.A + (size_t)r * g.lda + pc;
          if (pc + 3 < g.P) v = *reinterpret_cast<const float4*>(src);
          else { if (pc < g.P) v.x = src[0]; if (pc + 1 < g.P) v.y = src[1]; if (pc + 2 < g.P) v.z = src[2]; }
        }
        return v;
      };
      ra0 = ld(ra, p0 + c); ra1 = ld(rb2, p0 + c);
    } else {               // tile rows = p, 16 contiguous r
      const int lrow = tid >> 2, lk = (tid & 3) * 4;
      const int r0p = p0 + lrow, r1p = p0 + 64 + lrow;
      ra0 = r0p < g.P ? *reinterpret_cast<const float4*>(g.A + (size_t)r0p * g.lda + r0 + lk) : make_float4(0, 0, 0, 0);
      ra1 = r1p < g.P ? *reinterpret_cast<const float4*>(g.A + (size_t)r1p * g.lda + r0 + lk) : make_float4(0, 0, 0, 0);
    }
    if (B_RED_ROW) {       // 16 x 64 floats = 256 float4, 1 per thread
      const int rr = tid >> 4, c = (tid & 15) * 4;
      const int r = r0 + rr, qc = q0 + c;
      rb = make_float4(0, 0, 0, 0);
      if (r < g.R) {
        const float* src = g.B + (size_t)r * g.ldb + qc;
        if (qc + 3 < g.Q) rb = *reinterpret_cast<const float4*>(src);
        else { if (qc < g.Q) rb.x = src[0]; if (qc + 1 < g.Q) rb.y = src[1]; if (qc + 2 < g.Q) rb.z = src[2]; }
      }
    } else {
      const int lrow = tid >> 2, lk = (tid & 3) * 4;
      const int rq = q0 + lrow;
      rb = rq < g.Q ? *reinterpret_cast<const float4*>(g.B + (size_t)rq * g.ldb + r0 + lk) : make_float4(0, 0, 0, 0);
    }
  };
  auto sstore = [&](int buf) {
    if (A_RED_ROW) {
      const int rr = tid >> 5, c = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&As[buf][rr][c]) = ra0;
      *reinterpret_cast<float4*>(&As[buf][8 + rr][c]) = ra1;
    } else {
      const int lrow = tid >> 2, lk = (tid & 3) * 4;
      As[buf][lk + 0][lrow] = ra0.x; As[buf][lk + 1][lrow] = ra0.y; As[buf][lk + 2][lrow] = ra0.z; As[buf][lk + 3][lrow] = ra0.w;
      As[buf][lk + 0][64 + lrow] = ra1.x; As[buf][lk + 1][64 + lrow] = ra1.y; As[buf][lk + 2][64 + lrow] = ra1.z; As[buf][lk + 3][64 + lrow] = ra1.w;
    }
    if (B_RED_ROW) {
      const int rr = tid >> 4, c = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&Bs[buf][rr][c]) = rb;
    } else {
      const int lrow = tid >> 2, lk = (tid & 3) * 4;
      Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y; Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nk = (g.R + 15) / 16;
  gload(0); sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload((kb + 1) * 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
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
    if (kb + 1 < nk) { sstore(buf ^ 1); __syncthreads(); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int p = p0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (p >= g.P) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx * 4 + j;
      if (q >= g.Q) continue;
      float* c = g.C + (size_t)p * g.ldc + q;
      *c = g.accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

// strided naive GEMM for tiny dimensions (action_emb 7 -> d, action_pred d -> 7): C(i,j) = sum_r A(i,r) B(r,j) (+ bias_j)
struct NaiveArgs { const float* A; long sa_i, sa_r; const float* B; long sb_r, sb_j; const float* bias; float* C; long sc_i, sc_j; int I, J, R; };
__global__ void naive_gemm_kernel(NaiveArgs g) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)g.I * g.J) return;
  const int i = (int)(idx / g.J), j = (int)(idx % g.J);
  float acc = 0.f;
  for (int r = 0; r < g.R; ++r) acc = fmaf(g.A[i * g.sa_i + r * g.sa_r], g.B[r * g.sb_r + j * g.sb_j], acc);
  g.C[i * g.sc_i + j * g.sc_j] = acc + (g.bias ? g.bias[j] : 0.f);
}

// out[g, c] = sum_{t < T} src[(g*T + t), c]  (T = M: column sum).  One thread per output, fixed order -> deterministic.
__global__ void group_sum_kernel(const float* __restrict__ src, float* __restrict__ out, int G, int T, int C, int accumulate) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)G * C) return;
  const int gi = (int)(idx / C), c = (int)(idx % C);
  const float* p = src + (size_t)gi * T * C + c;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int t = 0;
  for (; t + 3 < T; t += 4) { s0 += p[(size_t)t * C]; s1 += p[(size_t)(t + 1) * C]; s2 += p[(size_t)(t + 2) * C]; s3 += p[(size_t)(t + 3) * C]; }
  for (; t < T; ++t) s0 += p[(size_t)t * C];
  const float s = (s0 + s1) + (s2 + s3);
  out[idx] = accumulate ? out[idx] + s : s;
}
// column sum over many rows in two deterministic stages: partial[blockIdx.y, c] over a row slab, then group_sum over slabs
__global__ void colsum_partial_kernel(const float* __restrict__ src, float* __restrict__ partial, int M, int C, int rows_per_slab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(M, r0 + rows_per_slab);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += src[(size_t)r * C + c];
  partial[(size_t)blockIdx.y * C + c] = s;
}

// activations: forward y = f(x) and backward dx = dy * f'(x) evaluated on the saved pre-activation
enum Act : int { ACT_GELU = 1, ACT_MISH = 2, ACT_SILU = 3 };
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long n, int act) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  y[i] = act == ACT_GELU ? gelu_erf(v) : act == ACT_MISH ? mish(v) : silu(v);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, long n, int act) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  float d;
  if (act == ACT_GELU) {
    d = 0.5f * (1.0f + erff(v * 0.70710678118654752440f)) + v * 0.39894228040143267794f * expf(-0.5f * v * v);
  } else if (act == ACT_SILU) {
    const float s = 1.0f / (1.0f + expf(-v));
    d = s * (1.0f + v * (1.0f - s));
  } else {   // mish: x tanh(softplus(x)) ; d = tanh(sp) + x (1 - tanh^2(sp)) sigmoid(x)
    const float sp = v > 20.0f ? v : log1pf(expf(v));
    const float th = tanhf(sp);
    const float s = 1.0f / (1.0f + expf(-v));
    d = th + v * (1.0f - th * th) * s;
  }
  dx[i] = dy[i] * d;
}

// fp32 [R, C] -> split-bf16 TRANSPOSED [C, 2*Rp] (hi | lo at column offset Rp), rows R..Rp zero-filled: turns the
// "reduce over rows" operands of dgrad (W^T) and wgrad (dy^T, x^T) into the K-major layout of the tensor-core GEMM.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int R, int C, int Rp) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? src[(size_t)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < C && r < Rp) {
      __nv_bfloat16 hi, lo;
      split_bf16(tile[tx][i], hi, lo);
      dst[(size_t)c * 2 * Rp + r] = hi;
      dst[(size_t)c * 2 * Rp + Rp + r] = lo;
    }
  }
}

// inverted dropout with a regenerable mask: out = x * [u(seed, i) >= p] / (1 - p)   (forward on x, backward on dy)
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ out, long n, float p, unsigned long long seed) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = x[i] * dropout_scale(seed, (unsigned long long)i, p);
}

// out = x + gate[g] * f  (gate may be null: out = x + f), g = row / rows_per_group
__global__ void gate_res_fwd_kernel(const float* __restrict__ x, const float* __restrict__ f, const float* __restrict__ gate,
                                    float* __restrict__ out, int M, int d, int rows_per_group) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)M * d) return;
  const int row = (int)(i / d), c = (int)(i % d);
  const float gv = gate ? gate[(size_t)(row / rows_per_group) * d + c] : 1.0f;
  out[i] = x[i] + gv * f[i];
}
// df = gate * dout ; prod = dout * f (its group sum is dgate)
__global__ void gate_res_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ f, const float* __restrict__ gate,
                                    float* __restrict__ df, float* __restrict__ prod, int M, int d, int rows_per_group) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)M * d) return;
  const int row = (int)(i / d), c = (int)(i % d);
  const float go = dout[i];
  df[i] = gate ? gate[(size_t)(row / rows_per_group) * d + c] * go : go;
  if (prod) prod[i] = go * f[i];
}

// LayerNorm(+modulate) backward, one warp per row.  Forward: n = xhat * w + b ; y = shift + n * scale (or y = n).
//   dn = dy * scale ; dx = rstd (gw - mean(gw) - xhat mean(gw xhat)), gw = dn * w
//   per-row terms for the parameter reductions: t_dw = dn * xhat, t_db = dn, t_dsc = dy * n  (dshift reduces dy itself)
struct LnBwdArgs {
  const float* x; const float* dy; const float* w; const float* b; const float* scale; int mod_stride; int rows_per_group;
  float* dx; float* t_dw; float* t_db; float* t_dsc; int M, d;
};
template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(LnBwdArgs a) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= a.M) return;
  const float* xr = a.x + (size_t)row * a.d;
  const float* dyr = a.dy + (size_t)row * a.d;
  float4 v[VPL], g[VPL];
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
  const size_t mrow = (size_t)(row / a.rows_per_group) * a.mod_stride;
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 w = *reinterpret_cast<const float4*>(a.w + c);
    const float4 dy4 = *reinterpret_cast<const float4*>(dyr + c);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (a.scale) sc = *reinterpret_cast<const float4*>(a.scale + mrow + c);
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.b) bb = *reinterpret_cast<const float4*>(a.b + c);
    const float xh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
    const float dyv[4] = {dy4.x, dy4.y, dy4.z, dy4.w}, scv[4] = {sc.x, sc.y, sc.z, sc.w}, wv[4] = {w.x, w.y, w.z, w.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
    float dn[4], gw[4], tdsc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dn[j] = dyv[j] * scv[j];
      gw[j] = dn[j] * wv[j];
      tdsc[j] = dyv[j] * (xh[j] * wv[j] + bv[j]);
      sg += gw[j]; sgx += gw[j] * xh[j];
    }
    g[i] = make_float4(gw[0], gw[1], gw[2], gw[3]);
    v[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
    const size_t o = (size_t)row * a.d + c;
    if (a.t_dw) *reinterpret_cast<float4*>(a.t_dw + o) = make_float4(dn[0] * xh[0], dn[1] * xh[1], dn[2] * xh[2], dn[3] * xh[3]);
    if (a.t_db) *reinterpret_cast<float4*>(a.t_db + o) = make_float4(dn[0], dn[1], dn[2], dn[3]);
    if (a.t_dsc) *reinterpret_cast<float4*>(a.t_dsc + o) = make_float4(tdsc[0], tdsc[1], tdsc[2], tdsc[3]);
  }
  const float mg = warp_sum(sg) / (float)a.d, mgx = warp_sum(sgx) / (float)a.d;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(a.dx + (size_t)row * a.d + c) =
        make_float4(rstd * (g[i].x - mg - v[i].x * mgx), rstd * (g[i].y - mg - v[i].y * mgx),
                    rstd * (g[i].z - mg - v[i].z * mgx), rstd * (g[i].w - mg - v[i].w * mgx));
  }
}

// Attention backward for tiny sequences, one CTA per (sample, head): recomputes P = softmax(q k^T scale + mask), then
//   dV = P^T dY ; dP = dY V^T ; dS = P (dP - rowsum(dP P)) ; dQ = dS K scale ; dK = dS^T Q scale      (transformer_blocks.py:119-158)
struct AttnBwdArgs {
  const float* q; int ldq; const float* k; const float* v; int ldkv; const float* dy; int lddy;
  float* dq; int lddq; float* dk; float* dv; int lddkv;
  int B, H, hd, Tq, Tk, causal; float scale;
  float p_drop; unsigned long long seed;     // dropout on the probabilities: same mask as the forward kernel
};
__global__ void __launch_bounds__(128) attention_bwd_kernel(AttnBwdArgs a) {
  __shared__ float sq[ATT_MAXT][ATT_MAXHD + 1], sk[ATT_MAXT][ATT_MAXHD + 1], sv[ATT_MAXT][ATT_MAXHD + 1], sdy[ATT_MAXT][ATT_MAXHD + 1];
  __shared__ float sp[ATT_MAXT][ATT_MAXT + 1], sds[ATT_MAXT][ATT_MAXT + 1], spu[ATT_MAXT][ATT_MAXT + 1];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, tid = threadIdx.x;
  const int hd = a.hd, Tq = a.Tq, Tk = a.Tk;
  for (int e = tid; e < Tq * hd; e += blockDim.x) {
    const int i = e / hd, c = e % hd;
    sq[i][c] = a.q[(size_t)(b * Tq + i) * a.ldq + h * hd + c];
    sdy[i][c] = a.dy[(size_t)(b * Tq + i) * a.lddy + h * hd + c];
  }
  for (int e = tid; e < Tk * hd; e += blockDim.x) {
    const int j = e / hd, c = e % hd;
    sk[j][c] = a.k[(size_t)(b * Tk + j) * a.ldkv + h * hd + c];
    sv[j][c] = a.v[(size_t)(b * Tk + j) * a.ldkv + h * hd + c];
  }
  __syncthreads();
  for (int e = tid; e < Tq * Tk; e += blockDim.x) {
    const int i = e / Tk, j = e % Tk;
    float s = 0.f, dp = 0.f;
    for (int c = 0; c < hd; ++c) { s = fmaf(sq[i][c], sk[j][c], s); dp = fmaf(sdy[i][c], sv[j][c], dp); }
    sp[i][j] = (a.causal && j > i) ? -INFINITY : s * a.scale;
    sds[i][j] = dp;
  }
  __syncthreads();
  if (tid < Tq) {
    float mx = -INFINITY;
    for (int j = 0; j < Tk; ++j) mx = fmaxf(mx, sp[tid][j]);
    float sum = 0.f;
    for (int j = 0; j < Tk; ++j) { float ex = expf(sp[tid][j] - mx); sp[tid][j] = ex; sum += ex; }
    const float inv = 1.0f / sum;
    float dot = 0.f;
    const unsigned long long base = ((unsigned long long)b * a.H * Tq + (unsigned long long)h * Tq + tid) * Tk;
    for (int j = 0; j < Tk; ++j) {
      sp[tid][j] *= inv;
      const float m = a.p_drop > 0.f ? dropout_scale(a.seed, base + j, a.p_drop) : 1.0f;   // P_used = P * m
      spu[tid][j] = sp[tid][j] * m;
      sds[tid][j] *= m;                         // dP = dP_used * m
      dot += sds[tid][j] * sp[tid][j];
    }
    for (int j = 0; j < Tk; ++j) sds[tid][j] = sp[tid][j] * (sds[tid][j] - dot) * a.scale;   // dS (scaled)
  }
  __syncthreads();
  for (int e = tid; e < Tq * hd; e += blockDim.x) {          // dQ = dS K
    const int i = e / hd, c = e % hd;
    float o = 0.f;
    for (int j = 0; j < Tk; ++j) o = fmaf(sds[i][j], sk[j][c], o);
    a.dq[(size_t)(b * Tq + i) * a.lddq + h * hd + c] = o;
  }
  for (int e = tid; e < Tk * hd; e += blockDim.x) {          // dK = dS^T Q ; dV = P^T dY
    const int j = e / hd, c = e % hd;
    float ok = 0.f, ov = 0.f;
    for (int i = 0; i < Tq; ++i) { ok = fmaf(sds[i][j], sq[i][c], ok); ov = fmaf(spu[i][j], sdy[i][c], ov); }
    a.dk[(size_t)(b * Tk + j) * a.lddkv + h * hd + c] = ok;
    a.dv[(size_t)(b * Tk + j) * a.lddkv + h * hd + c] = ov;
  }
}

// ------------------------------------------------------------------------------------------
// Fused multi-tensor AdamW + EMA (training step tail).  torch.optim.AdamW's single-tensor update (decoupled weight decay,
// bias-corrected moments) and the reference's EMA callback (mdt/callbacks/ema.py:117-126: ema -= (1 - decay) (ema - w)) for ALL
// parameter tensors of a group in ONE launch: the host passes a table of {param, grad, exp_avg, exp_avg_sq, ema, numel} and a
// block -> (tensor, chunk) map; every block updates one 4096-element chunk.
struct AdamTensor { float* p; const float* g; float* m; float* v; float* ema; long long n; float step_size; float bc2_sqrt; };   // 56 bytes
struct AdamHyper { float lr, beta1, beta2, eps, weight_decay, step_size, bc2_sqrt, ema_decay; int has_ema; const int* step_dev; };   // step_size / bc2_sqrt: per tensor
constexpr int ADAM_CHUNK = 4096;

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float* ema, const AdamHyper& h) {
  p *= 1.0f - h.lr * h.weight_decay;                    // param.mul_(1 - lr * wd)
  m = m + (1.0f - h.beta1) * (g - m);                   // exp_avg.lerp_(grad, 1 - beta1)
  v = v * h.beta2 + (1.0f - h.beta2) * g * g;           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / h.bc2_sqrt + h.eps;
  p = p - h.step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
  if (ema) { float e = *ema; e -= (1.0f - h.ema_decay) * (e - p); *ema = e; }
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(const AdamTensor* __restrict__ tab, const int2* __restrict__ blocks, AdamHyper h) {
  const int2 bc = blocks[blockIdx.x];
  const AdamTensor t = tab[bc.x];
  h.step_size = t.step_size; h.bc2_sqrt = t.bc2_sqrt;      // torch.optim.AdamW counts steps per parameter
  if (h.step_dev != nullptr) {                             // capturable mode: the step count lives on the device (CUDA-graph replays)
    const float tt = (float)*h.step_dev;
    h.step_size = h.lr / (1.0f - powf(h.beta1, tt));
    h.bc2_sqrt = sqrtf(1.0f - powf(h.beta2, tt));
  }
  const long long base = (long long)bc.y * ADAM_CHUNK;
  const long long end = base + ADAM_CHUNK < t.n ? base + ADAM_CHUNK : t.n;
  const bool vec = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) | reinterpret_cast<uintptr_t>(t.m) |
                     reinterpret_cast<uintptr_t>(t.v) | (h.has_ema ? reinterpret_cast<uintptr_t>(t.ema) : 0)) & 15) == 0;
  if (vec) {
    const long long end4 = base + ((end - base) & ~3LL);
    for (long long i = base + threadIdx.x * 4; i < end4; i += 256 * 4) {
      float4 p = *reinterpret_cast<float4*>(t.p + i), m = *reinterpret_cast<float4*>(t.m + i), v = *reinterpret_cast<float4*>(t.v + i);
      const float4 g = *reinterpret_cast<const float4*>(t.g + i);
      float4 e = h.has_ema ? *reinterpret_cast<float4*>(t.ema + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      adam_update(p.x, g.x, m.x, v.x, h.has_ema ? &e.x : nullptr, h); adam_update(p.y, g.y, m.y, v.y, h.has_ema ? &e.y : nullptr, h);
      adam_update(p.z, g.z, m.z, v.z, h.has_ema ? &e.z : nullptr, h); adam_update(p.w, g.w, m.w, v.w, h.has_ema ? &e.w : nullptr, h);
      *reinterpret_cast<float4*>(t.p + i) = p; *reinterpret_cast<float4*>(t.m + i) = m; *reinterpret_cast<float4*>(t.v + i) = v;
      if (h.has_ema) *reinterpret_cast<float4*>(t.ema + i) = e;
    }
    for (long long i = end4 + threadIdx.x; i < end; i += 256) adam_update(t.p[i], t.g[i], t.m[i], t.v[i], h.has_ema ? t.ema + i : nullptr, h);
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) adam_update(t.p[i], t.g[i], t.m[i], t.v[i], h.has_ema ? t.ema + i : nullptr, h);
  }
}

}  // namespace mdt
