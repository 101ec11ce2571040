// C-ABI of the second-generation training primitives (include/mdtb200.h, "training primitives, fused"): operand-emitting
// kernels (kernels_train2.cuh) and the tcgen05 GEMM on pre-split operands in its three roles.  Composed by
// mdt_policy_b200/training_fused.py into one autograd node per residual branch.  Included by engine.cu.
#pragma once
#include "kernels_train2.cuh"
#include "ops_train.cuh"

extern "C" {

// out16[M, 2K] = hi|lo split of x[M, K] (times act'(h) when h != NULL: the activation backward fused into the split);
// partial (optional): [ceil(M / 32), K] column sums of the same values -- mdtb200_op_group_sum(partial, out, 1, slabs, K) is the bias gradient
MDTB200_API int mdtb200_op_split(const float* x, const float* h, int act, void* out16, float* partial, int M, int K, void* stream) {
  if (!x || !out16 || M < 1 || K < 4 || K % 4 || (h && (act < ACT_GELU || act > ACT_SILU))) return op_fail(MDTB200_EINVAL, "op_split: bad argument");
  SplitArgs a{x, h, act, static_cast<__nv_bfloat16*>(out16), partial, M, K};
  split_rows_kernel<<<dim3((K + 511) / 512, (M + SPLIT_ROWS - 1) / SPLIT_ROWS), 128, 0, (cudaStream_t)stream>>>(a);
  return op_check("split_rows_kernel");
}
MDTB200_API int mdtb200_op_split_rows_per_slab(void) { return SPLIT_ROWS; }

// all weight operands of a step in one launch: table = device array of 32-byte records {const float* src, void* dst, int64 n, int32 K,
// int32 pad} (K > 0: dst = bf16 [n / K, 2K] hi|lo; K == 0: fp32 copy), blocks = device int2 {record, 4096-element chunk}
MDTB200_API int mdtb200_op_split_multi(const void* table, const void* blocks, int n_blocks, void* stream) {
  if (!table || !blocks || n_blocks < 1) return op_fail(MDTB200_EINVAL, "op_split_multi: bad argument");
  split_multi_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const SplitTensor*>(table), static_cast<const int2*>(blocks));
  return op_check("split_multi_kernel");
}

// GEMM on pre-split operands (bf16x3 on tcgen05, fp32 accumulate / output):
//   mode 0  C[M,N] = A[M,K] . B[N,K]^T + bias     A16 = x16 [M,2K], B16 = w16 [N,2K]; epi 6 (GELU16): C = pre-activation, C16 [M,2N] = split(GELU(C))
//   mode 1  C[M,K] = A[M,N] . B[N,K]              A16 = dy16 [M,2N], B16 = w16 [N,2K] read MN-major; epi 7 (GELUBWD16): C16 [M,2K] =
//           split(acc * GELU'(h)) with h passed in the bias slot and the [ceil(M/128), K] column partials of it in C (no fp32 output)
//   mode 2  C[N,K] = A[M,N]^T . B[M,K]            A16 = dy16 [M,2N], B16 = x16 [M,2K], both MN-major; splits > 1: deterministic split-K
//           over M (also accepted by mode 1, over N) with workspace sk_ws (splits * ceil(N/128) * ceil(K/64|128) * 128 * BN floats; mdtb200_op_gemm16_ws) and zeroed counters sk_cnt
MDTB200_API int64_t mdtb200_op_gemm16_ws(int N, int K, int splits) {
  return (int64_t)splits * ((N + 127) / 128) * ((K + 63) / 64) * 128 * 64;      // tiles x 128 x BN, BN | K-extent: same for BN = 64 / 128 / 192
}
MDTB200_API int mdtb200_op_gemm16(int mode, const void* A16, const void* B16, const float* bias, float* C, void* C16, int M, int N, int K,
                                  int epi, int splits, float* sk_ws, unsigned* sk_cnt, void* stream) {
  if (!A16 || !B16 || (!C && epi != EPI_GELUBWD16) || M < 1 || N < 1 || K < 1 || mode < 0 || mode > 2) return op_fail(MDTB200_EINVAL, "op_gemm16: bad argument");
  if (const char* e = op_tc_init()) return op_fail(MDTB200_ECUDA, "op_gemm16: %s", e);
  tc::TcGemm t{};
  t.A16 = static_cast<const __nv_bfloat16*>(A16); t.W16 = static_cast<const __nv_bfloat16*>(B16);
  t.bias = bias; t.C = C; t.rows_per_group = 1; t.epi = EPI_NONE; t.passes = 3; t.w_dynamic = 1;
  if (mode == 0) {
    if (K % 64 || N % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm16: forward needs K, N multiples of 64");
    if (epi != EPI_NONE && epi != EPI_GELU16) return op_fail(MDTB200_EINVAL, "op_gemm16: epilogue %d", epi);
    if (epi == EPI_GELU16 && !C16) return op_fail(MDTB200_EINVAL, "op_gemm16: GELU16 needs C16");
    t.lda16 = 2 * K; t.ldc = N; t.M = M; t.N = N; t.K = K; t.epi = epi;
    if (C16) { t.C16 = static_cast<__nv_bfloat16*>(C16); t.ldc16 = 2 * N; t.lo_off = N; }
  } else if (mode == 1) {
    if (N % 64 || K % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm16: dgrad needs N, K multiples of 64");
    t.lda16 = 2 * N; t.w_mn = 1; t.ldw16 = 2 * K; t.ldc = K; t.M = M; t.N = K; t.K = N;
    if (epi == EPI_GELUBWD16) {
      // activation backward fused into the epilogue: `bias` carries the pre-activation h [M, K] of the forward, `C` the [ceil(M/128), K]
      // column partials (bias gradient; may be NULL), C16 [M, 2K] receives split(acc * GELU'(h)); nothing is written in fp32
      if (!bias || !C16 || splits > 1) return op_fail(MDTB200_EINVAL, "op_gemm16: GELUBWD16 needs h (bias slot), C16 and no split-K");
      t.epi = EPI_GELUBWD16; t.R = bias; t.ldr = K; t.bias = nullptr; t.colpart = C; t.C = nullptr;
      t.C16 = static_cast<__nv_bfloat16*>(C16); t.ldc16 = 2 * K; t.lo_off = K;
    } else if (epi != EPI_NONE) return op_fail(MDTB200_EINVAL, "op_gemm16: epilogue %d", epi);
    if (splits > 1) { t.splits = splits; t.sk_ws = sk_ws; t.sk_cnt = sk_cnt; }
  } else {
    if (N % 64 || K % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm16: wgrad needs N, K multiples of 64");
    t.a_mn = 1; t.lda16 = 2 * N; t.w_mn = 1; t.ldw16 = 2 * K; t.ldc = K; t.M = N; t.N = K; t.K = M;
    if (splits > 1) { t.splits = splits; t.sk_ws = sk_ws; t.sk_cnt = sk_cnt; }
  }
  // 128-wide tiles whenever they divide the output: the training GEMMs have M >= 1536 rows, so 64-wide tiles only add waves and
  // operand re-reads (tools/sweep_train_tiles.sh: 5.84 -> 5.48 ms per graph-replayed step)
  t.bn_hint = ((t.M + 127) / 128) * (t.N / 128) * (t.splits > 1 ? t.splits : 1) >= 74 ? 128 : 64;      // ... unless that leaves half the SMs idle
  if (g_op_tma.cache.size() > 4096) g_op_tma.cache.clear();
  const char* e = tc::launch_tc_gemm(g_op_tma, t, (cudaStream_t)stream);
  if (e) return op_fail(MDTB200_ECUDA, "op_gemm16: %s", e);
  return op_check("tc_gemm_kernel (op16)");
}

// LayerNorm(+modulate) forward with optional fp32 (y) and split-bf16 (y16 [M, 2d]) outputs
MDTB200_API int mdtb200_op_ln_fwd16(const float* x, const float* w, const float* b, const float* shift, const float* scale, int mod_stride,
                                    int rows_per_group, int M, int d, float* y, void* y16, void* stream) {
  if (!x || !w || (!y && !y16) || M < 1 || d % 128 != 0 || d > 1024 || rows_per_group < 1) return op_fail(MDTB200_EINVAL, "op_ln_fwd16: bad argument");
  LnArgs a{};
  a.x = x; a.out = y; a.out16 = static_cast<__nv_bfloat16*>(y16); a.ld16 = 2 * d; a.lo_off = d;
  a.w = w; a.b = b; a.shift = shift; a.scale = scale; a.mod_stride = mod_stride; a.rows_per_group = rows_per_group; a.M = M; a.d = d;
  const int blocks = (M * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d / 128) {
    case 1: ln_mod_kernel<1><<<blocks, 256, 0, st>>>(a); break;
    case 2: ln_mod_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: ln_mod_kernel<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: ln_mod_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    default: return op_fail(MDTB200_EUNSUPPORTED, "op_ln_fwd16: d = %d", d);
  }
  return op_check("ln_mod_kernel (16)");
}

// LayerNorm(+modulate) backward per group of T rows, fused with the residual gradient (dx = dres + ...); dshift / dscale rows
// (stride dmod_stride) may be NULL; partial = [mdtb200_op_ln_bwd2_partials(M, T), 2d] (d ln.weight | d ln.bias), finished by mdtb200_op_group_sum
MDTB200_API int mdtb200_op_ln_bwd2_partials(int M, int T) {
  if (T <= 16 && M % T == 0) { const int R = T * (16 / T); return (M + R - 1) / R; }
  return ((M + T - 1) / T + 7) / 8;
}
MDTB200_API int mdtb200_op_ln_bwd2(const float* x, const float* dy, const float* w, const float* b, const float* scale, int mod_stride,
                                   const float* dres, float* dx, float* dshift, float* dscale, int dmod_stride, float* partial, int M, int d,
                                   int T, void* stream) {
  if (!x || !dy || !w || !dx || !partial || M < 1 || T < 1 || d % 128 != 0 || d > 512) return op_fail(MDTB200_EINVAL, "op_ln_bwd2: bad argument");
  LnBwd2Args a{x, dy, w, b, scale, mod_stride, dres, dx, dshift, dscale, dmod_stride, partial, M, d, T};
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 16 && M % T == 0) {      // row-parallel kernel: CTA = R rows = whole groups
    const int R = T * (16 / T), blocks = (M + R - 1) / R;
    const size_t smem = (size_t)R * 2 * d * sizeof(float);
    static bool configured = false;
    if (!configured) {
      cudaFuncSetAttribute(ln_bwd3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 128 * 4);
      cudaFuncSetAttribute(ln_bwd3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 256 * 4);
      cudaFuncSetAttribute(ln_bwd3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 384 * 4);
      cudaFuncSetAttribute(ln_bwd3_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 2 * 512 * 4);
      configured = true;
    }
    switch (d / 128) {
      case 1: ln_bwd3_kernel<1><<<blocks, 32 * R, smem, st>>>(a, R); break;
      case 2: ln_bwd3_kernel<2><<<blocks, 32 * R, smem, st>>>(a, R); break;
      case 3: ln_bwd3_kernel<3><<<blocks, 32 * R, smem, st>>>(a, R); break;
      case 4: ln_bwd3_kernel<4><<<blocks, 32 * R, smem, st>>>(a, R); break;
    }
    return op_check("ln_bwd3_kernel");
  }
  const int G = (M + T - 1) / T, blocks = (G + 7) / 8;
  switch (d / 128) {
    case 1: ln_bwd2_kernel<1><<<blocks, 256, 0, st>>>(a); break;
    case 2: ln_bwd2_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: ln_bwd2_kernel<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: ln_bwd2_kernel<4><<<blocks, 256, 0, st>>>(a); break;
  }
  return op_check("ln_bwd2_kernel");
}

// attention forward writing the split-bf16 operand of c_proj directly (y16 [B*Tq, 2D]); same kernel and dropout stream as mdtb200_op_attn_fwd
MDTB200_API int mdtb200_op_attn_fwd16(const float* q, int ldq, const float* k, const float* v, int ldkv, void* y16, int B, int H, int hd,
                                      int Tq, int Tk, int causal, float p_drop, uint64_t seed, void* stream) {
  if (!q || !k || !v || !y16 || B < 1 || H < 1 || hd < 4 || hd % 4 || hd > ATT_MAXHD || Tq < 1 || Tq > ATT_MAXT || Tk < 1 || Tk > ATT_MAXT ||
      !(p_drop >= 0.f && p_drop < 1.f))
    return op_fail(MDTB200_EINVAL, "op_attn_fwd16: bad argument");
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.k = k; a.v = v; a.ldkv = ldkv; a.y16 = static_cast<__nv_bfloat16*>(y16); a.ld16 = 2 * H * hd; a.lo_off = H * hd;
  a.B = B; a.H = H; a.hd = hd; a.Tq = Tq; a.Tk = Tk; a.causal = causal; a.scale = 1.0f / sqrtf((float)hd); a.p_drop = p_drop; a.seed = seed;
  // the shipped shapes run the compile-time specialised kernel (2 heads per CTA), as the sampling engine does
  const bool c = causal != 0;
  #define ATT_CASE(HD, TQ, TK, CA)                                                                                  \
    if (hd == HD && Tq == TQ && Tk == TK && c == (CA != 0) && H % 2 == 0) {                                         \
      attention_fixed_kernel<HD, TQ, TK, CA, 2><<<dim3(B, H / 2), 128, 0, (cudaStream_t)stream>>>(a);               \
      return op_check("attention_fixed_kernel (16)");                                                               \
    }
  ATT_CASE(48, 10, 10, 1) ATT_CASE(48, 10, 4, 1) ATT_CASE(48, 4, 4, 0)
  ATT_CASE(64, 10, 10, 1) ATT_CASE(64, 10, 3, 1) ATT_CASE(64, 3, 3, 0)
  #undef ATT_CASE
  const size_t smem = attention_smem_bytes(H * hd, H, Tq, Tk);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return op_fail(MDTB200_ECUDA, "attention smem %zu", smem);
    configured = smem;
  }
  attention_kernel<<<B, ATT_THREADS, smem, (cudaStream_t)stream>>>(a);
  return op_check("attention_kernel (16)");
}

// attention backward emitting operands (see attention_bwd2_kernel): shipped shapes only (MDTB200_EUNSUPPORTED otherwise: use mdtb200_op_attn_bwd
// + mdtb200_op_split).  Self-attention: dq16 = dkv16 = the (M, 2*3D) operand of the fused q|k|v projection (col0 = 0, D, 2D), bpart = (B, 3D).
// Cross-attention: dq16 = (M, 2D) operand of the query projection, dk32 / dv32 = fp32 halves of the (B*Tk, 2D) context gradient.
MDTB200_API int mdtb200_op_attn_bwd16(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* dy, int lddy,
                                      float* dk32, float* dv32, int ldkv32, void* dq16, int wq, int col0q, void* dkv16, int wkv, int col0k,
                                      int col0v, float* bpart, int bpart_kv, int B, int H, int hd, int Tq, int Tk, int causal, float p_drop,
                                      uint64_t seed, void* stream) {
  if (!q || !k || !v || !dy || !dq16 || B < 1 || H < 1 || !(p_drop >= 0.f && p_drop < 1.f)) return op_fail(MDTB200_EINVAL, "op_attn_bwd16: bad argument");
  AttnBwd2Args a{q, ldq, k, v, ldkv, dy, lddy, nullptr, 0, dk32, dv32, ldkv32, static_cast<__nv_bfloat16*>(dq16), wq, col0q,
                 static_cast<__nv_bfloat16*>(dkv16), wkv, col0k, col0v, bpart, bpart_kv, B, H, 1.0f / sqrtf((float)hd), p_drop, seed};
  const bool c = causal != 0;
  cudaStream_t st = (cudaStream_t)stream;
  #define ATTB_CASE(HD, TQ, TK, CA)                                                             \
    if (hd == HD && Tq == TQ && Tk == TK && c == (CA != 0)) {                                   \
      attention_bwd2_kernel<HD, TQ, TK, CA><<<B * H, 128, 0, st>>>(a);                          \
      return op_check("attention_bwd2_kernel");                                                 \
    }
  ATTB_CASE(48, 10, 10, 1) ATTB_CASE(48, 10, 4, 1) ATTB_CASE(48, 4, 4, 0)
  ATTB_CASE(64, 10, 10, 1) ATTB_CASE(64, 10, 3, 1) ATTB_CASE(64, 3, 3, 0)
  #undef ATTB_CASE
  return op_fail(MDTB200_EUNSUPPORTED, "op_attn_bwd16: shape (hd %d, Tq %d, Tk %d, causal %d) is not specialised", hd, Tq, Tk, causal);
}

// out = x + gate[row / T] * dropout(f; p, seed)        (gate NULL: 1; p == 0: identity mask)
MDTB200_API int mdtb200_op_res_drop_fwd(const float* x, const float* f, const float* gate, int gate_stride, float* out, int M, int d, int T,
                                        float p, uint64_t seed, void* stream) {
  if (!x || !f || !out || M < 1 || d < 4 || d % 4 || T < 1 || !(p >= 0.f && p < 1.f)) return op_fail(MDTB200_EINVAL, "op_res_drop_fwd: bad argument");
  ResFwdArgs a{x, f, gate, gate_stride, out, M, d, T, p, seed};
  const long n4 = (long)M * d / 4;
  res_drop_fwd_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return op_check("res_drop_fwd_kernel");
}
// backward: df16 [M, 2d] = split(gate * mask * dout); dgate rows (stride dgate_stride, NULL without gate) = sum_t dout * mask * f;
// bpartial (optional) [ceil(M / T), d] = per-group sums of df (bias gradient of the projection that produced f)
MDTB200_API int mdtb200_op_res_drop_bwd(const float* dout, const float* f, const float* gate, int gate_stride, void* df16, float* dgate,
                                        int dgate_stride, float* bpartial, int M, int d, int T, float p, uint64_t seed, void* stream) {
  if (!dout || !df16 || (dgate && !f) || M < 1 || d < 4 || d % 4 || T < 1 || !(p >= 0.f && p < 1.f)) return op_fail(MDTB200_EINVAL, "op_res_drop_bwd: bad argument");
  ResBwdArgs a{dout, f, gate, gate_stride, static_cast<__nv_bfloat16*>(df16), dgate, dgate_stride, bpartial, M, d, T, p, seed};
  res_drop_bwd_kernel<<<dim3((d + 511) / 512, (M + T - 1) / T), 128, 0, (cudaStream_t)stream>>>(a);
  return op_check("res_drop_bwd_kernel");
}

// y[M, J] = x[M, K] W[J, K]^T + bias, J <= 8 (output head)
MDTB200_API int mdtb200_op_narrow_fwd(const float* x, const float* W, const float* bias, float* y, int M, int K, int J, void* stream) {
  if (!x || !W || !y || M < 1 || K % 4 || J < 1 || J > 8) return op_fail(MDTB200_EINVAL, "op_narrow_fwd: bad argument");
  NarrowFwdArgs a{x, W, bias, y, M, K, J};
  narrow_out_kernel<<<(unsigned)(((long)M * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  return op_check("narrow_out_kernel");
}
// partial[ceil(M / 64)][N*J] of sum_m wide[m, n] thin[m, j] (layout [n, j] if wide_major else [j, n]); group_sum over the slabs finishes it
MDTB200_API int mdtb200_op_narrow_wgrad(const float* wide, const float* thin, float* partial, int M, int N, int J, int wide_major, void* stream) {
  if (!wide || !thin || !partial || M < 1 || N < 1 || J < 1 || J > 8) return op_fail(MDTB200_EINVAL, "op_narrow_wgrad: bad argument");
  NarrowWgradArgs a{wide, thin, partial, M, N, J, wide_major};
  narrow_wgrad_kernel<<<dim3((N + 127) / 128, (M + 63) / 64), 128, 0, (cudaStream_t)stream>>>(a);
  return op_check("narrow_wgrad_kernel");
}

}  // extern "C"
