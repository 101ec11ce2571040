// Stateless fp32 primitive ops of the TRAINING path, exported through the C ABI (include/mdtb200.h, "training primitives").
// The Python side (mdt_policy_b200/training.py) composes them into autograd Functions that mirror the reference's
// train-mode forward (score_wrappers.py:45-63 -> mdtv_transformer.py:208-236); every FLOP of forward and backward runs in
// these kernels.  Included by engine.cu (single translation unit).
#pragma once
#include "kernels_train2.cuh"

namespace {

int op_fail(int code, const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_create_error = buf;
  return code;
}
int op_check(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); return op_fail(MDTB200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e)); }
  return 0;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// process-wide TMA descriptor encoder / kernel configuration for the handle-less tensor-core ops
tc::TmaEncoder g_op_tma;
bool g_op_tc_ready = false;
const char* op_tc_init() {
  if (g_op_tc_ready) return nullptr;
  cudaDeviceProp prop{};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  if (prop.major != 10) return "tensor-core ops need an sm_100 device";
  if (const char* e = g_op_tma.init()) return e;
  if (const char* e = tc::configure_kernels()) return e;
  g_op_tc_ready = true;
  return nullptr;
}

}  // namespace

extern "C" {

// mode 0: C[M,N] = A[M,K] . B[N,K]^T (+ bias[N])     forward of nn.Linear (A = x, B = weight)
// mode 1: C[M,K] = A[M,N] . B[N,K]                  input gradient      (A = dy, B = weight)
// mode 2: C[N,K] (+)= A[M,N]^T . B[M,K]             weight gradient     (A = dy, B = x)
MDTB200_API int mdtb200_op_gemm(int mode, const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                int accumulate, void* stream) {
  if (!A || !B || !C || M < 1 || N < 1 || K < 1 || mode < 0 || mode > 2) return op_fail(MDTB200_EINVAL, "op_gemm: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = aligned16(A) && aligned16(B) && aligned16(C);
  if (mode == 0 && al && K % 16 == 0 && N % 4 == 0 && !accumulate && (!bias || aligned16(bias))) {
    GemmArgs g{};
    g.A = A; g.lda = K; g.W = B; g.bias = bias; g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K; g.rows_per_group = 1;
    dim3 grid((N + SG_BN - 1) / SG_BN, (M + SG_BM - 1) / SG_BM);
    sgemm_tn_kernel<EPI_NONE><<<grid, SG_THREADS, 0, st>>>(g);
    return op_check("sgemm_tn_kernel");
  }
  if (mode == 1 && al && N % 16 == 0 && K % 4 == 0 && !bias) {
    GGemmArgs g{A, N, B, K, C, K, M, K, N, accumulate};
    dim3 grid((K + 63) / 64, (M + 127) / 128);
    ggemm_kernel<false, true><<<grid, 256, 0, st>>>(g);
    return op_check("ggemm_kernel<dgrad>");
  }
  if (mode == 2 && al && N % 4 == 0 && K % 4 == 0 && !bias) {
    GGemmArgs g{A, N, B, K, C, K, N, K, M, accumulate};
    dim3 grid((K + 63) / 64, (N + 127) / 128);
    ggemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
    return op_check("ggemm_kernel<wgrad>");
  }
  if (accumulate) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm: accumulate needs 4-aligned dims");
  // tiny / odd dimensions (the 7-wide action embedding and output head): strided naive kernel
  NaiveArgs n{};
  n.bias = bias; n.C = C;
  if (mode == 0) { n.A = A; n.sa_i = K; n.sa_r = 1; n.B = B; n.sb_r = 1; n.sb_j = K; n.I = M; n.J = N; n.R = K; n.sc_i = N; n.sc_j = 1; }
  if (mode == 1) { n.A = A; n.sa_i = N; n.sa_r = 1; n.B = B; n.sb_r = K; n.sb_j = 1; n.I = M; n.J = K; n.R = N; n.sc_i = K; n.sc_j = 1; }
  if (mode == 2) { n.A = A; n.sa_i = 1; n.sa_r = N; n.B = B; n.sb_r = K; n.sb_j = 1; n.I = N; n.J = K; n.R = M; n.sc_i = K; n.sc_j = 1; }
  const long tot = (long)n.I * n.J;
  naive_gemm_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(n);
  return op_check("naive_gemm_kernel");
}

// Same three products on the tcgen05 tensor cores with split-bf16 (bf16x3) operands; fp32 in, fp32 out.
//   scratch: bf16 workspace of mdtb200_op_gemm_tc_scratch(mode, M, N, K) elements (operand copies in K-major hi|lo form:
//   forward splits x and W; dgrad splits dy and W^T; wgrad splits dy^T and x^T with the row count padded to 64).
// Requirements: reduce dim and output column count multiples of 64 (else MDTB200_EUNSUPPORTED: use mdtb200_op_gemm).
MDTB200_API int64_t mdtb200_op_gemm_tc_scratch(int mode, int M, int N, int K) {
  const int64_t Mp = (M + 127) / 128 * 128, M64 = (M + 63) / 64 * 64, Np = (N + 127) / 128 * 128;
  if (mode == 0) return Mp * 2 * K + (int64_t)N * 2 * K;
  if (mode == 1) return Mp * 2 * N + (int64_t)K * 2 * N;
  return Np * 2 * M64 + (int64_t)K * 2 * M64;
}

MDTB200_API int mdtb200_op_gemm_tc(int mode, const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                   void* scratch, void* stream) {
  if (!A || !B || !C || !scratch || M < 1 || N < 1 || K < 1 || mode < 0 || mode > 2) return op_fail(MDTB200_EINVAL, "op_gemm_tc: bad argument");
  if (const char* e = op_tc_init()) return op_fail(MDTB200_ECUDA, "op_gemm_tc: %s", e);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* s16 = reinterpret_cast<__nv_bfloat16*>(scratch);
  tc::TcGemm t{};
  t.bias = bias; t.C = C; t.rows_per_group = 1; t.epi = EPI_NONE; t.passes = 3;
  t.w_dynamic = 1;      // both operands are produced by the split kernels right before the GEMM on this stream
  const int64_t Mp = (M + 127) / 128 * 128;
  if (mode == 0) {          // y[M,N] = x[M,K] W[N,K]^T
    if (K % 64 || N % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm_tc: forward needs K, N multiples of 64");
    __nv_bfloat16 *a16 = s16, *w16 = s16 + Mp * 2 * K;
    split_weights_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256, 0, st>>>(A, a16, M, K);
    split_weights_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256, 0, st>>>(B, w16, N, K);
    t.A16 = a16; t.lda16 = 2 * K; t.W16 = w16; t.ldc = N; t.M = M; t.N = N; t.K = K;
  } else if (mode == 1) {   // dx[M,K] = dy[M,N] W[N,K] = dy . (W^T)[K,N]^T
    if (N % 64 || K % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm_tc: dgrad needs N, K multiples of 64");
    __nv_bfloat16 *a16 = s16, *w16 = s16 + Mp * 2 * N;
    split_weights_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256, 0, st>>>(A, a16, M, N);
    split_transpose_kernel<<<dim3((K + 31) / 32, (N + 31) / 32), 256, 0, st>>>(B, w16, N, K, N);     // W [N,K] -> [K, 2N]
    t.A16 = a16; t.lda16 = 2 * N; t.W16 = w16; t.ldc = K; t.M = M; t.N = K; t.K = N;
  } else {                  // dW[N,K] = dy[M,N]^T x[M,K] = (dy^T)[N,M] . (x^T)[K,M]^T, reduce dim M padded to 64
    if (K % 64) return op_fail(MDTB200_EUNSUPPORTED, "op_gemm_tc: wgrad needs K multiple of 64");
    const int M64 = (M + 63) / 64 * 64;
    const int64_t Np = (N + 127) / 128 * 128;
    __nv_bfloat16 *a16 = s16, *w16 = s16 + Np * 2 * M64;
    split_transpose_kernel<<<dim3((N + 31) / 32, (M64 + 31) / 32), 256, 0, st>>>(A, a16, M, N, M64);   // dy [M,N] -> [N, 2*M64]
    split_transpose_kernel<<<dim3((K + 31) / 32, (M64 + 31) / 32), 256, 0, st>>>(B, w16, M, K, M64);   // x  [M,K] -> [K, 2*M64]
    t.A16 = a16; t.lda16 = 2 * M64; t.W16 = w16; t.ldc = K; t.M = N; t.N = K; t.K = M64;
  }
  if (g_op_tma.cache.size() > 2048) g_op_tma.cache.clear();
  // launch_tc_gemm derives the W map's leading dimension from K: W16 is always [rows, 2K] contiguous here
  const char* e = tc::launch_tc_gemm(g_op_tma, t, st);
  if (e) return op_fail(MDTB200_ECUDA, "op_gemm_tc: %s", e);
  return op_check("tc_gemm_kernel (op)");
}

// out[g, c] (+)= sum_t src[g*T + t, c]
MDTB200_API int mdtb200_op_group_sum(const float* src, float* out, int G, int T, int Cn, int accumulate, void* stream) {
  if (!src || !out || G < 1 || T < 1 || Cn < 1) return op_fail(MDTB200_EINVAL, "op_group_sum: bad argument");
  launch_group_sum2(src, out, G, T, Cn, G * T, accumulate, (cudaStream_t)stream);
  return op_check("group_sum2_kernel");
}

// column sum of a tall matrix in two deterministic stages; `scratch` holds ceil(M / 64) * C floats
MDTB200_API int mdtb200_op_colsum(const float* src, float* out, float* scratch, int M, int Cn, int accumulate, void* stream) {
  if (!src || !out || !scratch || M < 1 || Cn < 1) return op_fail(MDTB200_EINVAL, "op_colsum: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int slabs = (M + 63) / 64;
  launch_group_sum2(src, scratch, slabs, 64, Cn, M, 0, st);
  launch_group_sum2(scratch, out, 1, slabs, Cn, slabs, accumulate, st);
  return op_check("colsum kernels");
}

// dy == NULL: out = act(x) ; else out = dy * act'(x).   act: 1 GELU(erf), 2 Mish, 3 SiLU
MDTB200_API int mdtb200_op_act(const float* x, const float* dy, float* out, int64_t n, int act, void* stream) {
  if (!x || !out || n < 1 || act < ACT_GELU || act > ACT_SILU) return op_fail(MDTB200_EINVAL, "op_act: bad argument");
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (dy) act_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, dy, out, (long)n, act);
  else act_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, out, (long)n, act);
  return op_check("act kernel");
}

// y = shift + LN(x; w, b) * scale   (shift/scale NULL: plain LayerNorm); shift/scale rows = (row / rows_per_group) * mod_stride
MDTB200_API int mdtb200_op_ln_fwd(const float* x, const float* w, const float* b, const float* shift, const float* scale, int mod_stride,
                                  int rows_per_group, int M, int d, float* y, void* stream) {
  if (!x || !w || !y || M < 1 || d % 128 != 0 || d > 1024 || rows_per_group < 1) return op_fail(MDTB200_EINVAL, "op_ln_fwd: bad argument");
  LnArgs a{};
  a.x = x; a.out = y; a.w = w; a.b = b; a.shift = shift; a.scale = scale; a.mod_stride = mod_stride; a.rows_per_group = rows_per_group; a.M = M; a.d = d;
  const int blocks = (M * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d / 128) {
    case 1: ln_mod_kernel<1><<<blocks, 256, 0, st>>>(a); break;
    case 2: ln_mod_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: ln_mod_kernel<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: ln_mod_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    case 6: ln_mod_kernel<6><<<blocks, 256, 0, st>>>(a); break;
    case 8: ln_mod_kernel<8><<<blocks, 256, 0, st>>>(a); break;
    default: return op_fail(MDTB200_EUNSUPPORTED, "op_ln_fwd: d = %d", d);
  }
  return op_check("ln_mod_kernel");
}

// backward of the op above: dx, and the per-row terms whose column / group sums are the parameter gradients:
//   t_dw = dn * xhat (-> d ln.weight), t_db = dn (-> d ln.bias), t_dsc = dy * n (-> d scale, group sum); d shift = group sum of dy
MDTB200_API int mdtb200_op_ln_bwd(const float* x, const float* dy, const float* w, const float* b, const float* scale, int mod_stride,
                                  int rows_per_group, int M, int d, float* dx, float* t_dw, float* t_db, float* t_dsc, void* stream) {
  if (!x || !dy || !w || !dx || M < 1 || d % 128 != 0 || d > 1024 || rows_per_group < 1) return op_fail(MDTB200_EINVAL, "op_ln_bwd: bad argument");
  LnBwdArgs a{x, dy, w, b, scale, mod_stride, rows_per_group, dx, t_dw, t_db, t_dsc, M, d};
  const int blocks = (M * 32 + 255) / 256;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d / 128) {
    case 1: ln_bwd_kernel<1><<<blocks, 256, 0, st>>>(a); break;
    case 2: ln_bwd_kernel<2><<<blocks, 256, 0, st>>>(a); break;
    case 3: ln_bwd_kernel<3><<<blocks, 256, 0, st>>>(a); break;
    case 4: ln_bwd_kernel<4><<<blocks, 256, 0, st>>>(a); break;
    case 6: ln_bwd_kernel<6><<<blocks, 256, 0, st>>>(a); break;
    case 8: ln_bwd_kernel<8><<<blocks, 256, 0, st>>>(a); break;
    default: return op_fail(MDTB200_EUNSUPPORTED, "op_ln_bwd: d = %d", d);
  }
  return op_check("ln_bwd_kernel");
}

// softmax(q k^T / sqrt(hd) + causal-top-left mask) v for T <= 16; strided operands (row stride in floats)
MDTB200_API int mdtb200_op_attn_fwd(const float* q, int ldq, const float* k, const float* v, int ldkv, float* y, int ldy, int B, int H, int hd,
                                    int Tq, int Tk, int causal, float p_drop, uint64_t seed, void* stream) {
  if (!q || !k || !v || !y || B < 1 || H < 1 || hd < 4 || hd % 4 || hd > ATT_MAXHD || Tq < 1 || Tq > ATT_MAXT || Tk < 1 || Tk > ATT_MAXT)
    return op_fail(MDTB200_EINVAL, "op_attn_fwd: bad argument");
  AttnArgs a{};
  a.q = q; a.ldq = ldq; a.k = k; a.v = v; a.ldkv = ldkv; a.y = y; a.ldy = ldy; a.B = B; a.H = H; a.hd = hd; a.Tq = Tq; a.Tk = Tk; a.causal = causal;
  a.scale = 1.0f / sqrtf((float)hd);
  a.p_drop = p_drop; a.seed = seed;
  if (!(p_drop >= 0.f && p_drop < 1.f)) return op_fail(MDTB200_EINVAL, "op_attn_fwd: p_drop must be in [0, 1)");
  const size_t smem = attention_smem_bytes(H * hd, H, Tq, Tk);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return op_fail(MDTB200_ECUDA, "attention smem %zu", smem);
    configured = smem;
  }
  attention_kernel<<<B, ATT_THREADS, smem, (cudaStream_t)stream>>>(a);
  return op_check("attention_kernel");
}

MDTB200_API int mdtb200_op_attn_bwd(const float* q, int ldq, const float* k, const float* v, int ldkv, const float* dy, int lddy, float* dq,
                                    int lddq, float* dk, float* dv, int lddkv, int B, int H, int hd, int Tq, int Tk, int causal, float p_drop,
                                    uint64_t seed, void* stream) {
  if (!q || !k || !v || !dy || !dq || !dk || !dv || B < 1 || H < 1 || hd < 1 || hd > ATT_MAXHD || Tq < 1 || Tq > ATT_MAXT || Tk < 1 || Tk > ATT_MAXT)
    return op_fail(MDTB200_EINVAL, "op_attn_bwd: bad argument");
  AttnBwdArgs a{q, ldq, k, v, ldkv, dy, lddy, dq, lddq, dk, dv, lddkv, B, H, hd, Tq, Tk, causal, 1.0f / sqrtf((float)hd), p_drop, seed};
  attention_bwd_kernel<<<B * H, 128, 0, (cudaStream_t)stream>>>(a);
  return op_check("attention_bwd_kernel");
}

// inverted dropout with a mask that is a pure function of (seed, element index): the same call on dy is the backward
MDTB200_API int mdtb200_op_dropout(const float* x, float* out, int64_t n, float p, uint64_t seed, void* stream) {
  if (!x || !out || n < 1 || !(p >= 0.f && p < 1.f)) return op_fail(MDTB200_EINVAL, "op_dropout: bad argument");
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, out, (long)n, p, seed);
  return op_check("dropout_kernel");
}

// out = x + gate[row / rows_per_group] * f   (gate NULL: out = x + f)
MDTB200_API int mdtb200_op_gate_res(const float* x, const float* f, const float* gate, float* out, int M, int d, int rows_per_group, void* stream) {
  if (!x || !f || !out || M < 1 || d < 1 || rows_per_group < 1) return op_fail(MDTB200_EINVAL, "op_gate_res: bad argument");
  const long n = (long)M * d;
  gate_res_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, f, gate, out, M, d, rows_per_group);
  return op_check("gate_res_fwd_kernel");
}
// df = gate * dout ; prod = dout * f (group-sum it for d gate; may be NULL)
MDTB200_API int mdtb200_op_gate_res_bwd(const float* dout, const float* f, const float* gate, float* df, float* prod, int M, int d,
                                        int rows_per_group, void* stream) {
  if (!dout || !f || !df || M < 1 || d < 1 || rows_per_group < 1) return op_fail(MDTB200_EINVAL, "op_gate_res_bwd: bad argument");
  const long n = (long)M * d;
  gate_res_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dout, f, gate, df, prod, M, d, rows_per_group);
  return op_check("gate_res_bwd_kernel");
}

// Fused multi-tensor AdamW (+ EMA) step: `table` = device array of AdamTensor records (56 bytes: param, grad, exp_avg, exp_avg_sq,
// ema pointers, int64 numel, float step_size = lr / (1 - beta1^t), float bc2_sqrt = sqrt(1 - beta2^t) with t the PER-PARAMETER step
// count, as torch.optim.AdamW keeps it), `blocks` = device array of n_blocks int2 {tensor index, 4096-element chunk index}.
// step_dev (optional): device int32 step count shared by all tensors of the call -- bias corrections are then computed in the
// kernel from it (capturable mode: the host-side values would be frozen inside a CUDA graph).
MDTB200_API int mdtb200_op_adamw_ema(const void* table, const void* blocks, int n_blocks, float lr, float beta1, float beta2, float eps,
                                     float weight_decay, float ema_decay, int has_ema, const int* step_dev, void* stream) {
  if (!table || !blocks || n_blocks < 1) return op_fail(MDTB200_EINVAL, "op_adamw_ema: bad argument");
  AdamHyper h{lr, beta1, beta2, eps, weight_decay, 0.f, 1.f, ema_decay, has_ema, step_dev};
  adamw_ema_kernel<<<n_blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const AdamTensor*>(table), static_cast<const int2*>(blocks), h);
  return op_check("adamw_ema_kernel");
}

// Registers a device counter that is mixed into every dropout seed (NULL: off).  Lets a CUDA-graph-captured training step draw
// fresh dropout masks on every replay: the captured graph increments the counter, the host-side seeds stay frozen.  Process-wide.
MDTB200_API int mdtb200_op_set_seed_epoch(const uint64_t* dev_counter) {
  const unsigned long long* p = reinterpret_cast<const unsigned long long*>(dev_counter);
  cudaError_t e = cudaMemcpyToSymbol(g_seed_epoch, &p, sizeof(p));
  if (e != cudaSuccess) return op_fail(MDTB200_ECUDA, "op_set_seed_epoch: %s", cudaGetErrorString(e));
  return 0;
}

}  // extern "C"
