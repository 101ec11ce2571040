// placeholder (stage A): tensor-core GEMM not wired yet
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
namespace mdt { namespace tc {
struct TmaEncoder { const char* init() { return "tcgen05 path not built"; } };
struct TcGemm {
  const __nv_bfloat16* A16; int lda16; const __nv_bfloat16* W16; const float* bias;
  float* C; int ldc; __nv_bfloat16* C16; int ldc16; int lo_off;
  const float* R; int ldr; const float* gate; int gate_stride; int rows_per_group;
  int M, N, K, epi, passes;
};
inline const char* configure_kernels() { return "tcgen05 path not built"; }
inline const char* launch_tc_gemm(TmaEncoder&, const TcGemm&, cudaStream_t) { return "tcgen05 path not built"; }
}}
