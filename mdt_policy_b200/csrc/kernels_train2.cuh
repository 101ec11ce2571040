// Training path, second generation: kernels that hand each other tensor-core operands.
//
// Every GEMM of the training step (forward, input gradient, weight gradient) runs on tc::tc_gemm_kernel from ONE row-major
// hi|lo bf16 split per tensor ([rows, 2 cols]): the forward reads x16 / W16 K-major, dgrad reads W16 MN-major, wgrad reads dy16
// and x16 MN-major (gemm_tcgen05.cuh), so no transposed copies exist.  The kernels here produce those splits as a by-product of
// the work they do anyway (LayerNorm forward, attention forward, residual backward, activation backward) and fold the small
// reductions (bias / LayerNorm / AdaLN gradients) into per-CTA partial sums that one group_sum launch finishes.
// Reference semantics: mdt/models/networks/transformers/transformer_blocks.py (Block :209-214, ConditionedBlock :292-309).
#pragma once
#include "kernels_train.cuh"

namespace mdt {

__device__ __forceinline__ float act_grad(float v, int act) {
  if (act == ACT_GELU) return gelu_erf_grad(v);
  if (act == ACT_SILU) { const float s = 1.0f / (1.0f + expf(-v)); return s * (1.0f + v * (1.0f - s)); }
  const float sp = v > 20.0f ? v : log1pf(expf(v));      // mish
  const float th = tanhf(sp);
  const float s = 1.0f / (1.0f + expf(-v));
  return th + v * (1.0f - th * th) * s;
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* dst, int lo_off, const float (&o)[4]) {
  __align__(8) __nv_bfloat16 hi[4];
  __align__(8) __nv_bfloat16 lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16(o[j], hi[j], lo[j]);
  *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(dst + lo_off) = *reinterpret_cast<const uint2*>(lo);
}

// out16[r, :] = split(x[r, :] * (h ? act'(h[r, :]) : 1)),  partial[slab, c] = column sums over the slab's rows (bias gradient).
// grid (ceil(K / 512), ceil(M / SPLIT_ROWS)), 128 threads, thread = 4 columns.
struct SplitArgs { const float* x; const float* h; int act; __nv_bfloat16* out16; float* partial; int M, K; };
constexpr int SPLIT_ROWS = 32;
__global__ void __launch_bounds__(128) split_rows_kernel(SplitArgs a) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
  if (c >= a.K) return;
  const int r0 = blockIdx.y * SPLIT_ROWS, r1 = min(a.M, r0 + SPLIT_ROWS);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float4 v = *reinterpret_cast<const float4*>(a.x + (size_t)r * a.K + c);
    float o[4] = {v.x, v.y, v.z, v.w};
    if (a.h) {
      const float4 hv = *reinterpret_cast<const float4*>(a.h + (size_t)r * a.K + c);
      o[0] *= act_grad(hv.x, a.act); o[1] *= act_grad(hv.y, a.act); o[2] *= act_grad(hv.z, a.act); o[3] *= act_grad(hv.w, a.act);
    }
    s[0] += o[0]; s[1] += o[1]; s[2] += o[2]; s[3] += o[3];
    store_split4(a.out16 + (size_t)r * 2 * a.K + c, a.K, o);
  }
  if (a.partial) *reinterpret_cast<float4*>(a.partial + (size_t)blockIdx.y * a.K + c) = make_float4(s[0], s[1], s[2], s[3]);
}

// One launch for all the weights of a step: table entry = {src fp32 [n / K, K], dst, n, K}; K > 0: dst = bf16 [n / K, 2K] hi|lo
// split (rows of grouped weights are laid out back to back, so q/k/v share one operand); K == 0: dst = fp32 copy (grouped biases).
struct SplitTensor { const float* src; void* dst; long long n; int K; int pad; };   // 32 bytes
constexpr int SPLITM_CHUNK = 4096;
__global__ void __launch_bounds__(256) split_multi_kernel(const SplitTensor* __restrict__ tab, const int2* __restrict__ blocks) {
  const int2 blk = blocks[blockIdx.x];
  const SplitTensor t = tab[blk.x];
  const long long base = (long long)blk.y * SPLITM_CHUNK;
#pragma unroll
  for (int u = 0; u < SPLITM_CHUNK / (256 * 4); ++u) {
    const long long i = base + (u * 256 + threadIdx.x) * 4;
    if (i >= t.n) continue;
    const float4 v = *reinterpret_cast<const float4*>(t.src + i);
    if (t.K == 0) { *reinterpret_cast<float4*>(static_cast<float*>(t.dst) + i) = v; continue; }
    const long long r = i / t.K;
    const int c = (int)(i - r * t.K);
    const float o[4] = {v.x, v.y, v.z, v.w};
    store_split4(static_cast<__nv_bfloat16*>(t.dst) + r * 2 * t.K + c, t.K, o);
  }
}

// out = x + gate[g] * dropout(f)     (gate null: out = x + dropout(f); p == 0: no dropout), g = row / T, one thread per float4
struct ResFwdArgs { const float* x; const float* f; const float* gate; int gate_stride; float* out; int M, d, T; float p; unsigned long long seed; };
__global__ void __launch_bounds__(256) res_drop_fwd_kernel(ResFwdArgs a) {
  const long i4 = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long i = i4 * 4;
  if (i >= (long)a.M * a.d) return;
  const int row = (int)(i / a.d), c = (int)(i % a.d);
  const float4 xv = *reinterpret_cast<const float4*>(a.x + i), fv = *reinterpret_cast<const float4*>(a.f + i);
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
  if (a.gate) g = *reinterpret_cast<const float4*>(a.gate + (size_t)(row / a.T) * a.gate_stride + c);
  float m[4] = {1.f, 1.f, 1.f, 1.f};
  if (a.p > 0.f) {
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = dropout_scale(a.seed, (unsigned long long)(i + j), a.p);
  }
  *reinterpret_cast<float4*>(a.out + i) = make_float4(xv.x + g.x * m[0] * fv.x, xv.y + g.y * m[1] * fv.y, xv.z + g.z * m[2] * fv.z, xv.w + g.w * m[3] * fv.w);
}
// backward of the op above, emitting the c_proj output gradient directly as a GEMM operand:
//   df16 = split(gate * mask * dout),  dgate[g, :] = sum_t dout * mask * f,  bpartial[g, :] = sum_t df  (c_proj bias gradient)
// thread = (group g, 4 columns), loops the T rows of the group.
struct ResBwdArgs {
  const float* dout; const float* f; const float* gate; int gate_stride; __nv_bfloat16* df16; float* dgate; int dgate_stride;
  float* bpartial; int M, d, T; float p; unsigned long long seed;
};
__global__ void __launch_bounds__(128) res_drop_bwd_kernel(ResBwdArgs a) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 4, g = blockIdx.y;
  if (c >= a.d) return;
  float4 gv = make_float4(1.f, 1.f, 1.f, 1.f);
  if (a.gate) gv = *reinterpret_cast<const float4*>(a.gate + (size_t)g * a.gate_stride + c);
  const float gt[4] = {gv.x, gv.y, gv.z, gv.w};
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, bs[4] = {0.f, 0.f, 0.f, 0.f};
  const int r1 = min(a.M, (g + 1) * a.T);
  for (int r = g * a.T; r < r1; ++r) {
    const size_t i = (size_t)r * a.d + c;
    const float4 dv = *reinterpret_cast<const float4*>(a.dout + i);
    const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
    float m[4] = {1.f, 1.f, 1.f, 1.f};
    if (a.p > 0.f) {
#pragma unroll
      for (int j = 0; j < 4; ++j) m[j] = dropout_scale(a.seed, (unsigned long long)(i + j), a.p);
    }
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { o[j] = gt[j] * m[j] * dd[j]; bs[j] += o[j]; }
    if (a.dgate) {
      const float4 fv = *reinterpret_cast<const float4*>(a.f + i);
      dg[0] = fmaf(dd[0] * m[0], fv.x, dg[0]); dg[1] = fmaf(dd[1] * m[1], fv.y, dg[1]);
      dg[2] = fmaf(dd[2] * m[2], fv.z, dg[2]); dg[3] = fmaf(dd[3] * m[3], fv.w, dg[3]);
    }
    store_split4(a.df16 + (size_t)r * 2 * a.d + c, a.d, o);
  }
  if (a.dgate) *reinterpret_cast<float4*>(a.dgate + (size_t)g * a.dgate_stride + c) = make_float4(dg[0], dg[1], dg[2], dg[3]);
  if (a.bpartial) *reinterpret_cast<float4*>(a.bpartial + (size_t)g * a.d + c) = make_float4(bs[0], bs[1], bs[2], bs[3]);
}

// LayerNorm(+modulate) backward, one warp per GROUP of T rows (a sample), fused with the residual-stream gradient:
//   dx = dres + rstd (gw - mean(gw) - xhat mean(gw xhat)),  dn = dy * scale, gw = dn * w            (dres may be null)
//   dshift[g] = sum_t dy, dscale[g] = sum_t dy * n    (written straight into the AdaLN gradient rows; null for a plain LN)
//   partial[cta, 0:d] = sum dn * xhat (-> d ln.weight), partial[cta, d:2d] = sum dn (-> d ln.bias)  over the CTA's rows
struct LnBwd2Args {
  const float* x; const float* dy; const float* w; const float* b; const float* scale; int mod_stride;
  const float* dres; float* dx; float* dshift; float* dscale; int dmod_stride; float* partial; int M, d, T;
};
template <int VPL>
__global__ void __launch_bounds__(256) ln_bwd2_kernel(LnBwd2Args a) {
  __shared__ float red[8][2 * VPL * 128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x * 8 + warp, G = (a.M + a.T - 1) / a.T;
  float4 adw[VPL], adb[VPL], ash[VPL], asc[VPL], wv4[VPL], bv4[VPL], sc4[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    adw[i] = adb[i] = ash[i] = asc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv4[i] = *reinterpret_cast<const float4*>(a.w + c);
    bv4[i] = a.b ? *reinterpret_cast<const float4*>(a.b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    sc4[i] = (a.scale && g < G) ? *reinterpret_cast<const float4*>(a.scale + (size_t)g * a.mod_stride + c) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (g < G) {
    const int r1 = min(a.M, (g + 1) * a.T);
    for (int row = g * a.T; row < r1; ++row) {
      const float* xr = a.x + (size_t)row * a.d;
      const float* dyr = a.dy + (size_t)row * a.d;
      float4 v[VPL], gq[VPL], dy4[VPL];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4);
        dy4[i] = *reinterpret_cast<const float4*>(dyr + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      const float mean = warp_sum(s) / (float)a.d;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
        q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)a.d + 1e-5f);
      float sg = 0.f, sgx = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const float xh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
        const float dyv[4] = {dy4[i].x, dy4[i].y, dy4[i].z, dy4[i].w}, scv[4] = {sc4[i].x, sc4[i].y, sc4[i].z, sc4[i].w};
        const float wv[4] = {wv4[i].x, wv4[i].y, wv4[i].z, wv4[i].w}, bv[4] = {bv4[i].x, bv4[i].y, bv4[i].z, bv4[i].w};
        float dn[4], gw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dn[j] = dyv[j] * scv[j];
          gw[j] = dn[j] * wv[j];
          sg += gw[j]; sgx += gw[j] * xh[j];
        }
        adw[i].x += dn[0] * xh[0]; adw[i].y += dn[1] * xh[1]; adw[i].z += dn[2] * xh[2]; adw[i].w += dn[3] * xh[3];
        adb[i].x += dn[0]; adb[i].y += dn[1]; adb[i].z += dn[2]; adb[i].w += dn[3];
        ash[i].x += dyv[0]; ash[i].y += dyv[1]; ash[i].z += dyv[2]; ash[i].w += dyv[3];
        asc[i].x += dyv[0] * (xh[0] * wv[0] + bv[0]); asc[i].y += dyv[1] * (xh[1] * wv[1] + bv[1]);
        asc[i].z += dyv[2] * (xh[2] * wv[2] + bv[2]); asc[i].w += dyv[3] * (xh[3] * wv[3] + bv[3]);
        gq[i] = make_float4(gw[0], gw[1], gw[2], gw[3]);
        v[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
      }
      const float mg = warp_sum(sg) / (float)a.d, mgx = warp_sum(sgx) / (float)a.d;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 o = make_float4(rstd * (gq[i].x - mg - v[i].x * mgx), rstd * (gq[i].y - mg - v[i].y * mgx),
                               rstd * (gq[i].z - mg - v[i].z * mgx), rstd * (gq[i].w - mg - v[i].w * mgx));
        if (a.dres) {
          const float4 dr = *reinterpret_cast<const float4*>(a.dres + (size_t)row * a.d + c);
          o.x += dr.x; o.y += dr.y; o.z += dr.z; o.w += dr.w;
        }
        *reinterpret_cast<float4*>(a.dx + (size_t)row * a.d + c) = o;
      }
    }
    if (a.dshift) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int c = (i * 32 + lane) * 4;
        *reinterpret_cast<float4*>(a.dshift + (size_t)g * a.dmod_stride + c) = ash[i];
        *reinterpret_cast<float4*>(a.dscale + (size_t)g * a.dmod_stride + c) = asc[i];
      }
    }
  }
  // CTA reduction of the affine-parameter partials in warp order (deterministic)
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(&red[warp][c]) = adw[i];
    *reinterpret_cast<float4*>(&red[warp][VPL * 128 + c]) = adb[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * VPL * 128; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][c];
    a.partial[(size_t)blockIdx.x * 2 * VPL * 128 + c] = s;
  }
}

// Row-parallel variant of the kernel above (one warp per ROW, CTA = R = T * groups-per-CTA <= 16 rows): the per-row terms of the
// parameter / modulation gradients meet in shared memory and are summed in row order (deterministic).  partial = [ceil(M / R), 2d].
template <int VPL>
__global__ void __launch_bounds__(512) ln_bwd3_kernel(LnBwd2Args a, int R) {
  extern __shared__ __align__(16) float lnb_smem[];            // [R][2d]
  constexpr int d = VPL * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * R + warp;
  const bool live = row < a.M;
  float4 t1[VPL], t2[VPL], u1[VPL], u2[VPL];                   // (dn xhat | dn) and (dy | dy n)
#pragma unroll
  for (int i = 0; i < VPL; ++i) t1[i] = t2[i] = u1[i] = u2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const int g = row / a.T;
    const float* xr = a.x + (size_t)row * d;
    const float* dyr = a.dy + (size_t)row * d;
    float4 v[VPL], dy4[VPL], wv4[VPL], bv4[VPL], sc4[VPL], dr4[VPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = *reinterpret_cast<const float4*>(xr + c);
      dy4[i] = *reinterpret_cast<const float4*>(dyr + c);
      wv4[i] = *reinterpret_cast<const float4*>(a.w + c);
      bv4[i] = a.b ? *reinterpret_cast<const float4*>(a.b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      sc4[i] = a.scale ? *reinterpret_cast<const float4*>(a.scale + (size_t)g * a.mod_stride + c) : make_float4(1.f, 1.f, 1.f, 1.f);
      dr4[i] = a.dres ? *reinterpret_cast<const float4*>(a.dres + (size_t)row * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / (float)d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
    float sg = 0.f, sgx = 0.f;
    float4 gq[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const float xh[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
      const float dyv[4] = {dy4[i].x, dy4[i].y, dy4[i].z, dy4[i].w}, scv[4] = {sc4[i].x, sc4[i].y, sc4[i].z, sc4[i].w};
      const float wv[4] = {wv4[i].x, wv4[i].y, wv4[i].z, wv4[i].w}, bv[4] = {bv4[i].x, bv4[i].y, bv4[i].z, bv4[i].w};
      float dn[4], gw[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dn[j] = dyv[j] * scv[j];
        gw[j] = dn[j] * wv[j];
        sg += gw[j]; sgx += gw[j] * xh[j];
      }
      t1[i] = make_float4(dn[0] * xh[0], dn[1] * xh[1], dn[2] * xh[2], dn[3] * xh[3]);
      t2[i] = make_float4(dn[0], dn[1], dn[2], dn[3]);
      u1[i] = dy4[i];
      u2[i] = make_float4(dyv[0] * (xh[0] * wv[0] + bv[0]), dyv[1] * (xh[1] * wv[1] + bv[1]), dyv[2] * (xh[2] * wv[2] + bv[2]), dyv[3] * (xh[3] * wv[3] + bv[3]));
      gq[i] = make_float4(gw[0], gw[1], gw[2], gw[3]);
      v[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
    }
    const float mg = warp_sum(sg) / (float)d, mgx = warp_sum(sgx) / (float)d;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(a.dx + (size_t)row * d + c) =
          make_float4(dr4[i].x + rstd * (gq[i].x - mg - v[i].x * mgx), dr4[i].y + rstd * (gq[i].y - mg - v[i].y * mgx),
                      dr4[i].z + rstd * (gq[i].z - mg - v[i].z * mgx), dr4[i].w + rstd * (gq[i].w - mg - v[i].w * mgx));
    }
  }
  float* mine = lnb_smem + (size_t)warp * 2 * d;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(mine + c) = t1[i];
    *reinterpret_cast<float4*>(mine + d + c) = t2[i];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * d; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += lnb_smem[(size_t)r * 2 * d + c];
    a.partial[(size_t)blockIdx.x * 2 * d + c] = s;
  }
  if (a.dshift) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int c = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(mine + c) = u1[i];
      *reinterpret_cast<float4*>(mine + d + c) = u2[i];
    }
    __syncthreads();
    const int gpc = R / a.T, G = a.M / a.T;
    for (int e = threadIdx.x; e < gpc * 2 * d; e += blockDim.x) {
      const int gl = e / (2 * d), c = e % (2 * d), g = blockIdx.x * gpc + gl;
      if (g >= G) continue;
      float s = 0.f;
      for (int t = 0; t < a.T; ++t) s += lnb_smem[(size_t)(gl * a.T + t) * 2 * d + c];
      if (c < d) a.dshift[(size_t)g * a.dmod_stride + c] = s;
      else a.dscale[(size_t)g * a.dmod_stride + c - d] = s;
    }
  }
}

// out[g, c] (+)= sum_{t < T, g*T + t < rows} src[(g*T + t), c]: block = 32 columns x RL row lanes, lanes summed in fixed order
template <int RL>
__global__ void __launch_bounds__(32 * RL) group_sum2_kernel(const float* __restrict__ src, float* __restrict__ out, int G, int T, int C, int rows, int accumulate) {
  __shared__ float red[RL][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx, g = blockIdx.y;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    const int t1 = min(T, rows - g * T);
    const float* p = src + (size_t)g * T * C + c;
    int t = ty;
    for (; t + 3 * RL < t1; t += 4 * RL) {
      s0 += p[(size_t)t * C]; s1 += p[(size_t)(t + RL) * C]; s2 += p[(size_t)(t + 2 * RL) * C]; s3 += p[(size_t)(t + 3 * RL) * C];
    }
    for (; t < t1; t += RL) s0 += p[(size_t)t * C];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < RL; ++r) s += red[r][tx];
    out[(size_t)g * C + c] = accumulate ? out[(size_t)g * C + c] + s : s;
  }
}
inline void launch_group_sum2(const float* src, float* out, int G, int T, int C, int rows, int accumulate, cudaStream_t st) {
  const dim3 grid((C + 31) / 32, G);
  if (T > 64) group_sum2_kernel<32><<<grid, 1024, 0, st>>>(src, out, G, T, C, rows, accumulate);      // long columns: 32 row lanes
  else group_sum2_kernel<8><<<grid, 256, 0, st>>>(src, out, G, T, C, rows, accumulate);
}

// Attention backward for the shipped tiny shapes, compile-time specialised (one CTA per (sample, head), 128 threads), emitting the
// projections' output gradients directly as GEMM operands:
//   P = softmax(q k^T scale + causal top-left mask), P_used = P * dropout mask (same hash stream as the forward kernels)
//   dV = P_used^T dY ; dP = (dY V^T) * mask ; dS = P (dP - rowsum(dP P)) scale ; dQ = dS K ; dK = dS^T Q     (transformer_blocks.py:119-158)
// dq goes to dq32 (fp32, row stride ldq32) and/or dq16 (bf16 hi|lo rows of width wq: element (row, col0q + h*HD + c), lo at + wq), the same
// for dk / dv with (ldkv32, wkv, col0k / col0v); bpart (optional) = [B, wq] per-sample column sums of dq (and of dk / dv when they share
// the dq16 operand, i.e. self-attention) -> one group_sum over the batch gives the q|k|v bias gradients.
struct AttnBwd2Args {
  const float* q; int ldq; const float* k; const float* v; int ldkv; const float* dy; int lddy;
  float* dq32; int ldq32; float* dk32; float* dv32; int ldkv32;
  __nv_bfloat16* dq16; int wq; int col0q; __nv_bfloat16* dkv16; int wkv; int col0k; int col0v;
  float* bpart; int bpart_kv;       // bpart_kv: 1 = dk / dv column sums go to bpart as well (columns col0k / col0v of the same wq-wide row)
  int B, H; float scale; float p_drop; unsigned long long seed;
};
template <int HD, int TQ, int TK, int CAUSAL>
__global__ void __launch_bounds__(128) attention_bwd2_kernel(AttnBwd2Args a) {
  constexpr int HP = HD + 4, H4 = HD / 4;
  __shared__ __align__(16) float sq[TQ * HP], sdy[TQ * HP], sk[TK * HP], sv[TK * HP];
  __shared__ __align__(16) float so[(TQ > TK ? TQ : TK) * HP];        // staging of an output block for the column sums
  __shared__ float sp[TQ][TK + 1], sds[TQ][TK + 1], spu[TQ][TK + 1];
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H, tid = threadIdx.x;
  for (int e = tid; e < TQ * H4; e += 128) {
    const int i = e / H4, c = (e % H4) * 4;
    *reinterpret_cast<float4*>(sq + i * HP + c) = *reinterpret_cast<const float4*>(a.q + (size_t)(b * TQ + i) * a.ldq + h * HD + c);
    *reinterpret_cast<float4*>(sdy + i * HP + c) = *reinterpret_cast<const float4*>(a.dy + (size_t)(b * TQ + i) * a.lddy + h * HD + c);
  }
  for (int e = tid; e < TK * H4; e += 128) {
    const int j = e / H4, c = (e % H4) * 4;
    *reinterpret_cast<float4*>(sk + j * HP + c) = *reinterpret_cast<const float4*>(a.k + (size_t)(b * TK + j) * a.ldkv + h * HD + c);
    *reinterpret_cast<float4*>(sv + j * HP + c) = *reinterpret_cast<const float4*>(a.v + (size_t)(b * TK + j) * a.ldkv + h * HD + c);
  }
  __syncthreads();
  for (int e = tid; e < TQ * TK; e += 128) {
    const int i = e / TK, j = e % TK;
    float s = 0.f, dp = 0.f;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(sq + i * HP + c), kv = *reinterpret_cast<const float4*>(sk + j * HP + c);
      const float4 gv = *reinterpret_cast<const float4*>(sdy + i * HP + c), vv = *reinterpret_cast<const float4*>(sv + j * HP + c);
      s = fmaf(qv.x, kv.x, s); s = fmaf(qv.y, kv.y, s); s = fmaf(qv.z, kv.z, s); s = fmaf(qv.w, kv.w, s);
      dp = fmaf(gv.x, vv.x, dp); dp = fmaf(gv.y, vv.y, dp); dp = fmaf(gv.z, vv.z, dp); dp = fmaf(gv.w, vv.w, dp);
    }
    sp[i][j] = (CAUSAL && j > i) ? -INFINITY : s * a.scale;
    sds[i][j] = dp;
  }
  __syncthreads();
  if (tid < TQ) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TK; ++j) mx = fmaxf(mx, sp[tid][j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < TK; ++j) { const float ex = expf(sp[tid][j] - mx); sp[tid][j] = ex; sum += ex; }
    const float inv = 1.0f / sum;
    float dot = 0.f;
    const unsigned long long base = ((unsigned long long)b * a.H * TQ + (unsigned long long)h * TQ + tid) * TK;
#pragma unroll
    for (int j = 0; j < TK; ++j) {
      sp[tid][j] *= inv;
      const float m = a.p_drop > 0.f ? dropout_scale(a.seed, base + j, a.p_drop) : 1.0f;
      spu[tid][j] = sp[tid][j] * m;
      sds[tid][j] *= m;
      dot += sds[tid][j] * sp[tid][j];
    }
#pragma unroll
    for (int j = 0; j < TK; ++j) sds[tid][j] = sp[tid][j] * (sds[tid][j] - dot) * a.scale;
  }
  __syncthreads();
  // one output block at a time (dQ, dK, dV): thread = (row, 4 columns); results to global (fp32 and / or split bf16) and to `so`
  auto emit = [&](int rows, int which) {       // which: 0 dQ, 1 dK, 2 dV
    for (int e = tid; e < rows * H4; e += 128) {
      const int r = e / H4, c = (e % H4) * 4;
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      if (which == 0) {
#pragma unroll
        for (int j = 0; j < TK; ++j) {
          const float w = sds[r][j];
          const float4 x = *reinterpret_cast<const float4*>(sk + j * HP + c);
          o[0] = fmaf(w, x.x, o[0]); o[1] = fmaf(w, x.y, o[1]); o[2] = fmaf(w, x.z, o[2]); o[3] = fmaf(w, x.w, o[3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < TQ; ++i) {
          const float w = which == 1 ? sds[i][r] : spu[i][r];
          const float4 x = *reinterpret_cast<const float4*>((which == 1 ? sq : sdy) + i * HP + c);
          o[0] = fmaf(w, x.x, o[0]); o[1] = fmaf(w, x.y, o[1]); o[2] = fmaf(w, x.z, o[2]); o[3] = fmaf(w, x.w, o[3]);
        }
      }
      *reinterpret_cast<float4*>(so + r * HP + c) = make_float4(o[0], o[1], o[2], o[3]);
      const size_t grow = (size_t)b * (which == 0 ? TQ : TK) + r;
      if (which == 0) {
        if (a.dq32) *reinterpret_cast<float4*>(a.dq32 + grow * a.ldq32 + h * HD + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (a.dq16) store_split4(a.dq16 + grow * 2 * a.wq + a.col0q + h * HD + c, a.wq, o);
      } else {
        float* d32 = which == 1 ? a.dk32 : a.dv32;
        if (d32) *reinterpret_cast<float4*>(d32 + grow * a.ldkv32 + h * HD + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (a.dkv16) store_split4(a.dkv16 + grow * 2 * a.wkv + (which == 1 ? a.col0k : a.col0v) + h * HD + c, a.wkv, o);
      }
    }
    if (a.bpart && (which == 0 || a.bpart_kv)) {
      __syncthreads();
      for (int c = tid; c < HD; c += 128) {
        float s = 0.f;
        for (int r = 0; r < rows; ++r) s += so[r * HP + c];
        a.bpart[(size_t)b * a.wq + (which == 0 ? a.col0q : which == 1 ? a.col0k : a.col0v) + h * HD + c] = s;
      }
      __syncthreads();
    }
  };
  emit(TQ, 0);
  emit(TK, 1);
  emit(TK, 2);
}

// Narrow linear layers (the 7-wide action embedding / output head), M rows:
//   forward  y[m, j] = sum_k x[m, k] W[j, k] + b[j]   (J <= 8 outputs, warp per row)
struct NarrowFwdArgs { const float* x; const float* W; const float* bias; float* y; int M, K, J; };
__global__ void __launch_bounds__(256) narrow_out_kernel(NarrowFwdArgs a) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= a.M) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int k = lane * 4; k < a.K; k += 128) {
    const float4 xv = *reinterpret_cast<const float4*>(a.x + (size_t)row * a.K + k);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < a.J) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.W + (size_t)j * a.K + k));
        acc[j] = fmaf(xv.x, w.x, acc[j]); acc[j] = fmaf(xv.y, w.y, acc[j]); acc[j] = fmaf(xv.z, w.z, acc[j]); acc[j] = fmaf(xv.w, w.w, acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j < a.J) {
      const float v = warp_sum(acc[j]);
      if (lane == 0) a.y[(size_t)row * a.J + j] = v + (a.bias ? a.bias[j] : 0.f);
    }
  }
}
//   weight gradient partials over 64-row slabs: P[slab][n, j] (wide_major) or P[slab][j, n] = sum_m wide[m, n] * thin[m, j]
//   (thin has J <= 8 columns; d action_emb.weight [N, J] = dy^T x : wide = dy, thin = x, wide_major;
//    d action_pred.weight [J, N] = dy^T x : wide = x, thin = dy).  A group_sum over the slabs finishes it.
struct NarrowWgradArgs { const float* wide; const float* thin; float* partial; int M, N, J, wide_major; };
__global__ void __launch_bounds__(128) narrow_wgrad_kernel(NarrowWgradArgs a) {
  __shared__ float st[64][8];
  const int n = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * 64, rows = min(64, a.M - r0);
  for (int e = threadIdx.x; e < 64 * 8; e += 128) {
    const int r = e >> 3, j = e & 7;
    st[r][j] = (r < rows && j < a.J) ? a.thin[(size_t)(r0 + r) * a.J + j] : 0.f;
  }
  __syncthreads();
  if (n >= a.N) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int r = 0; r < rows; ++r) {
    const float wv = a.wide[(size_t)(r0 + r) * a.N + n];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(wv, st[r][j], acc[j]);
  }
  float* P = a.partial + (size_t)blockIdx.y * a.N * a.J;
  for (int j = 0; j < a.J; ++j) P[a.wide_major ? (size_t)n * a.J + j : (size_t)j * a.N + n] = acc[j];
}

}  // namespace mdt
