// tcgen05 / TMEM / TMA GEMM for sm_100a with fused epilogues and split-bf16 ("bf16x3") operands.
//
//   D[M,N] = epi( A[M,K] . W[N,K]^T + bias )        A, W: fp32 values stored as bf16 hi | lo halves
//
// Both operands are K-major ([rows, 2K] bf16: columns [0,K) = hi, [K,2K) = lo), staged by TMA
// (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage shared-memory ring; one elected thread issues
// tcgen05.mma (cta_group::1, kind::f16, M=128, N=BN, K=16) with the fp32 accumulator in TMEM; with PASSES == 3
// each k-block issues hi*hi + lo*hi + hi*lo into the same accumulator (~2^-17 relative operand error, see
// tools/emulate_split_precision.py), with PASSES == 1 only hi*hi (plain bf16).  Warp roles (512 threads):
//   warp 0  TMA producer      warp 1  MMA issuer      warp 2  TMEM allocator      all 16 warps: epilogue
// Epilogue (TMEM -> registers via tcgen05.ld 32x32b): + bias, GELU(erf), residual, AdaLN gate, then fp32 store
// and/or a split-bf16 store that directly produces the next GEMM's A operand.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>
#include <map>
#include <tuple>

#include "kernels_simt.cuh"

namespace mdt { namespace tc {

constexpr int BM = 128;        // UMMA M (rows of A per CTA)
constexpr int BK = 64;         // bf16 elements per k-block = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int THREADS = 512;       // 16 warps: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator; all 16 run the epilogue

struct TcGemm {
  const __nv_bfloat16* A16; int lda16; const __nv_bfloat16* W16; const float* bias;
  float* C; int ldc; __nv_bfloat16* C16; int ldc16; int lo_off;
  const float* R; int ldr; const float* gate; int gate_stride; int rows_per_group;
  int M, N, K, epi, passes;
  unsigned long long* trace;
  int gi, go, goff;            // output-row remap (row = (m / gi) * go + goff + m % gi; gi == 0: identity), C / C16 only
  int wg_rows, wg_stride, w_rows;   // weight groups: rows [g * wg_rows, +wg_rows) of A use W rows [g * wg_stride + n, ...) (w_rows = total W rows)
  int w_dynamic;                    // 1: W is written by the preceding kernel on the stream (training ops): no W prefetch ahead of the PDL wait
  // MN-major operands (training: one row-major hi|lo split per tensor serves forward, dgrad and wgrad -- no transposed copies):
  //   a_mn: A16 is [K (reduce) rows, lda16] with element (m, k) at A16[k * lda16 + m], lo half at column lda16 / 2 + m
  //   w_mn: W16 is [K (reduce) rows, ldw16] with element (n, k) at W16[k * ldw16 + n], lo half at column ldw16 / 2 + n
  // K need not be a multiple of 64 when both operands are MN-major (TMA zero-fills the rows past K).
  int a_mn, w_mn, ldw16;
  // split-K (deterministic): grid.z = splits CTAs per tile write fp32 partial tiles to sk_ws, the last one to arrive (sk_cnt, self-
  // resetting counters, one per tile, zeroed once by the owner) adds them in slice order and runs the epilogue.
  int splits; float* sk_ws; unsigned* sk_cnt;
  float* colpart;                   // EPI_GELUBWD16: [ceil(M / 128), N] column sums of the emitted operand per row tile (bias gradient partials)
  int bn_hint;                      // 0: tile-width policy of launch_tc_gemm; 64 / 128 / 192: use this width when it divides N
};

struct TcParams {
  const float* bias;
  float* C; int ldc; __nv_bfloat16* C16; int ldc16; int lo_off;
  const float* R; int ldr; const float* gate; int gate_stride; int rows_per_group;
  int M, N, K, epi;
  unsigned long long* trace;   // optional (tests): per-CTA globaltimer stamps [cta][8]
  int gi, go, goff;
  int wg_rows, wg_stride;
  int w_early;                 // 1: W is static (inference weights): its first tiles may be requested before the dependency wait
  int a_mn, w_mn, a_lo, w_lo;  // MN-major operand flags and the column offset of their lo halves
  int splits; float* sk_ws; unsigned* sk_cnt;
  float* colpart;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(slot) do { if (p.trace) p.trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (slot)] = gtimer(); } while (0)

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must surface as a CUDA error (trap), never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 (16 B,
// ignored for swizzled K-major) | SBO = 1024 B (8 rows x 128 B) | version 1 (bits 46-47) | layout SWIZZLE_128B (2, bits 61-63)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major SWIZZLE_128B descriptor (cute mma_traits_sm100.hpp, make_umma_desc<Major::MN>: canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))
// in 16-byte units): the tile is stored as TMA boxes of 64 reduce-rows x 64 MN-elements (128-byte rows): 8 k-rows = 1024 B (SBO),
// the next 64 MN-elements are the next 8 KB box (LBO); one UMMA_K = 16 step advances the start address by 16 rows = 2048 B.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), K-major both (bit 15 / 16 = 1: A / B MN-major),
// N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int BN, int PASSES>
struct SmemLayout {
  static constexpr int A_TILE = BM * BK * 2;                 // 16 KB
  static constexpr int W_TILE = BN * BK * 2;
  static constexpr int STAGE = (PASSES == 3 ? 2 : 1) * (A_TILE + W_TILE);
  static constexpr int NT = THREADS;
  // one CTA per SM with the deepest ring that fits: the operand stream is bound by per-SM ingest (~100 GB/s/SM measured), and
  // 2-stage / 2-CTA-per-SM variants measured 9 % slower end to end (profiles/r01_experiments.md)
  static constexpr int STAGES = (PASSES == 3) ? (BN == 192 ? 2 : BN == 128 ? 3 : 4) : (BN == 192 ? 4 : BN == 128 ? 5 : 6);
  static constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);      // power of two >= BN
  static constexpr int MIN_CTAS = 1;
  static constexpr int BAR_OFF = STAGES * STAGE;
  static constexpr int BIAS_OFF = BAR_OFF + 256;             // BN floats: this tile's bias slice (staged in the prologue)
  static constexpr int TOTAL = BIAS_OFF + 1024 + 1024;       // barriers + tmem slot, bias, + slack for 1024-B alignment
};

// ------------------------------------------------------------------------------------------ the kernel
// EXT = true compiles in the training extensions (MN-major operands, split-K with in-kernel reduction, GELU16); the sampling engine
// runs the EXT = false instantiations, whose code is exactly the lean inference kernel.
template <int BN, int PASSES, bool EXT = false>
__global__ void __launch_bounds__((SmemLayout<BN, PASSES>::NT), (SmemLayout<BN, PASSES>::MIN_CTAS))
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  const bool a_mn = EXT && p.a_mn, w_mn = EXT && p.w_mn;
  const int splits = EXT ? p.splits : 1;
  using L = SmemLayout<BN, PASSES>;
  constexpr int NT = L::NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (L::STAGES + s); };
  const uint32_t tmem_full = bar0 + 8u * (2 * L::STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::BAR_OFF + 8 * (2 * L::STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int n0w = n0 + (p.wg_rows ? (m0 / p.wg_rows) * p.wg_stride : 0);     // W row of this tile (weight groups, e.g. one per head)
  const int nkb_all = EXT ? (p.K + BK - 1) / BK : p.K / BK;
  const int kb_beg = splits > 1 ? (int)((long)nkb_all * blockIdx.z / splits) : 0;
  const int kb_end = splits > 1 ? (int)((long)nkb_all * (blockIdx.z + 1) / splits) : nkb_all;
  const int nkb = kb_end - kb_beg;           // k-blocks of this CTA (split-K: a slice of the reduce dimension)
  if (threadIdx.x == 0) TC_STAMP(0);
  // operand tile loads: K-major = one box {64 k, rows}; MN-major = boxes {64 mn, 64 k} (8 KB each) along the MN extent
  auto load_a = [&](uint32_t dst, uint32_t bar, int kb, int lo) {
    if (!a_mn) tma_load_2d(dst, &tmA, bar, (lo ? p.K : 0) + kb * BK, m0);
    else {
#pragma unroll
      for (int j = 0; j < BM / 64; ++j) tma_load_2d(dst + j * 8192, &tmA, bar, (lo ? p.a_lo : 0) + m0 + 64 * j, kb * BK);
    }
  };
  auto load_w = [&](uint32_t dst, uint32_t bar, int kb, int lo) {
    if (!w_mn) tma_load_2d(dst, &tmW, bar, (lo ? p.K : 0) + kb * BK, n0w);
    else {
#pragma unroll
      for (int j = 0; j < BN / 64; ++j) tma_load_2d(dst + j * 8192, &tmW, bar, (lo ? p.w_lo : 0) + n0w + 64 * j, kb * BK);
    }
  };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < L::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), L::TMEM_COLS);
  // static weights (w_early) come with a static bias: its slice is staged in shared memory here, before the dependency wait, so the
  // epilogue does not pay an L2 round trip after the last MMA
  float* sbias = reinterpret_cast<float*>(smem + L::BIAS_OFF);
  const bool bias_staged = p.w_early && p.bias && splits <= 1;
  if (bias_staged && threadIdx.x >= 128 && threadIdx.x < 128 + BN) sbias[threadIdx.x - 128] = p.bias[n0 + threadIdx.x - 128];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // The weight tiles do not depend on the previous kernel: the producer requests them for the first ring pass BEFORE the
  // dependency wait, so that under PDL they stream in while the predecessor is still draining its epilogue.
  const int npre = p.w_early ? (nkb < L::STAGES ? nkb : L::STAGES) : 0;
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < npre; ++kb) {
      const uint32_t st = sbase + kb * L::STAGE;
      mbar_expect_tx(full_bar(kb), L::STAGE);
      load_w(st + L::A_TILE, full_bar(kb), kb_beg + kb, 0);                                         // W hi
      if (PASSES == 3) load_w(st + 2 * L::A_TILE + L::W_TILE, full_bar(kb), kb_beg + kb, 1);        // W lo
    }
  }
  pdl_wait(KT_GEMM);       // prologue above overlapped the previous kernel's tail; its results are visible from here on
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (kb / L::STAGES) & 1;
        const uint32_t st = sbase + s * L::STAGE;
        if (kb >= npre) {
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), L::STAGE);
          load_w(st + L::A_TILE, full_bar(s), kb_beg + kb, 0);                                         // W hi
          if (PASSES == 3) load_w(st + 2 * L::A_TILE + L::W_TILE, full_bar(s), kb_beg + kb, 1);        // W lo
        }
        load_a(st, full_bar(s), kb_beg + kb, 0);                                                       // A hi
        if (PASSES == 3) load_a(st + L::A_TILE + L::W_TILE, full_bar(s), kb_beg + kb, 1);              // A lo
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, BN) | (a_mn ? 1u << 15 : 0u) | (w_mn ? 1u << 16 : 0u);
      const uint32_t a_step = a_mn ? 2048 >> 4 : (UMMA_K * 2) >> 4, w_step = w_mn ? 2048 >> 4 : (UMMA_K * 2) >> 4;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % L::STAGES;
        const uint32_t ph = (kb / L::STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (kb == 0) TC_STAMP(2);
        if (kb == nkb - 1) TC_STAMP(3);
        const uint32_t st = sbase + s * L::STAGE;
        const uint64_t a_hi = a_mn ? make_smem_desc_mn(st) : make_smem_desc(st);
        const uint64_t w_hi = w_mn ? make_smem_desc_mn(st + L::A_TILE) : make_smem_desc(st + L::A_TILE);
        const uint64_t a_lo = a_mn ? make_smem_desc_mn(st + L::A_TILE + L::W_TILE) : make_smem_desc(st + L::A_TILE + L::W_TILE);
        const uint64_t w_lo = w_mn ? make_smem_desc_mn(st + 2 * L::A_TILE + L::W_TILE) : make_smem_desc(st + 2 * L::A_TILE + L::W_TILE);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: +32 bytes per UMMA_K inside the 128-byte swizzle row; MN-major: +16 rows of 128 bytes
          const uint64_t ada = (uint64_t)(k * a_step), adw = (uint64_t)(k * w_step);
          if (PASSES == 3) {
            umma_f16(tmem_base, a_lo + ada, w_hi + adw, idesc, (kb | k) != 0);   // small terms first
            umma_f16(tmem_base, a_hi + ada, w_lo + adw, idesc, 1);
            umma_f16(tmem_base, a_hi + ada, w_hi + adw, idesc, 1);
          } else {
            umma_f16(tmem_base, a_hi + ada, w_hi + adw, idesc, (kb | k) != 0);
          }
        }
        umma_commit(empty_bar(s));          // smem slot reusable once these MMAs have read it
      }
      umma_commit(tmem_full);               // accumulator complete
    }
  }
  __syncwarp();
  {
    // ===== epilogue: TMEM -> registers -> global, by ALL 16 warps =====
    // warp w may only touch TMEM lanes [32 (w % 4), +32): the four warps of a lane quarter split the BN columns.
    const int q = warp & 3, grp = warp >> 2;
    constexpr int CPW = BN / (NT / 128);            // columns per warp: 32 (BN = 128) or 16 (BN = 64)
    constexpr int GPR = BN / 8;                            // 8-column groups per row (phase B work items)
    constexpr int ITEMS = BM * GPR / NT;              // phase-B items per thread: 2 (BN = 64) or 4 (BN = 128)
    // residual / gate operands of phase B are fetched now, while the mainloop is still running (this hides their L2
    // latency, ~1 us); the BN = 64 and the lean BN = 128 instantiations do it (d-wide outputs), the 192-wide / training ones do not
    constexpr bool PRE = BN == 64 || (BN == 128 && !EXT);
    float4 pre_r[PRE ? ITEMS * 2 : 1], pre_g[PRE ? ITEMS * 2 : 1];
    const bool prefetched = PRE && (p.epi == EPI_RES || p.epi == EPI_RES_GATE) && warp >= 2;
    if (PRE && prefetched) {
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = threadIdx.x + it * NT;
        const int r = idx / GPR, cg = (idx % GPR) * 8, row = m0 + r;
        if (row < p.M) {
          const float* rr = p.R + (size_t)row * p.ldr + n0 + cg;
          pre_r[it * 2] = *reinterpret_cast<const float4*>(rr); pre_r[it * 2 + 1] = *reinterpret_cast<const float4*>(rr + 4);
          if (p.epi == EPI_RES_GATE) {
            const float* gp = p.gate + (size_t)(row / p.rows_per_group) * p.gate_stride + n0 + cg;
            pre_g[it * 2] = *reinterpret_cast<const float4*>(gp); pre_g[it * 2 + 1] = *reinterpret_cast<const float4*>(gp + 4);
          }
        }
      }
    }
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    pdl_trigger();    // all MMAs of this CTA are done: the next kernel may start its prologue while we drain TMEM
    if (threadIdx.x == 64) TC_STAMP(4);
    // phase A (thread = accumulator row): TMEM -> +bias -> activation -> fp32 staging tile in the (now idle) pipeline smem
    constexpr int SP = BN + 4;                             // padded row stride (floats) of the staging tile
    float* stage = reinterpret_cast<float*>(smem);
    {
      const int r = q * 32 + lane;
#pragma unroll
      for (int c = 0; c < CPW / 16; ++c) {
        const int col = grp * CPW + c * 16;
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
        if (EXT && nkb == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 0.f;        // empty split-K slice: nothing was accumulated
        }
        if (p.bias && splits <= 1) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 b = bias_staged ? *reinterpret_cast<const float4*>(sbias + col + j) : *reinterpret_cast<const float4*>(p.bias + n0 + col + j);
            o[j] += b.x; o[j + 1] += b.y; o[j + 2] += b.z; o[j + 3] += b.w;
          }
        }
        if (p.epi == EPI_GELU && splits <= 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = gelu_erf_fast(o[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(stage + r * SP + col + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) TC_STAMP(5);     // a warp-2 thread: phase A done (did not run the producer / MMA loops)
    __syncthreads();
    if (threadIdx.x == 64) TC_STAMP(6);
    bool finish = true;
    if (EXT && splits > 1) {
      // split-K: publish this slice's raw tile, count arrivals; the last CTA of the tile sums the slices in slice order
      // (deterministic), applies bias / activation and continues into phase B; the others are done.
      const int tile = blockIdx.y * gridDim.x + blockIdx.x;
      float* ws = p.sk_ws + (size_t)tile * splits * (BM * BN);
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = threadIdx.x + it * NT;
        const int r = idx / GPR, cg = (idx % GPR) * 8;
        float* dst = ws + (size_t)blockIdx.z * (BM * BN) + r * BN + cg;
        __stcg(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(stage + r * SP + cg));
        __stcg(reinterpret_cast<float4*>(dst + 4), *reinterpret_cast<const float4*>(stage + r * SP + cg + 4));
      }
      __threadfence();
      __syncthreads();
      volatile int* s_last = reinterpret_cast<volatile int*>(smem + L::BAR_OFF + 8 * (2 * L::STAGES + 1) + 8);
      if (threadIdx.x == 0) {
        const unsigned old = atomicAdd(p.sk_cnt + tile, 1u);
        const int last = old == (unsigned)(splits - 1);
        if (last) p.sk_cnt[tile] = 0u;              // self-reset: the next launch finds the counters zero again
        *s_last = last;
      }
      __syncthreads();
      finish = *s_last != 0;
      if (finish) {
        __threadfence();
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
          const int idx = threadIdx.x + it * NT;
          const int r = idx / GPR, cg = (idx % GPR) * 8;
          float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
          for (int z = 0; z < splits; ++z) {
            const float* src = ws + (size_t)z * (BM * BN) + r * BN + cg;
            const float4 a = __ldcg(reinterpret_cast<const float4*>(src)), b = __ldcg(reinterpret_cast<const float4*>(src + 4));
            s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
          }
          float o[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += p.bias[n0 + cg + j];
          }
          if (p.epi == EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = gelu_erf_fast(o[j]);
          }
          *reinterpret_cast<float4*>(stage + r * SP + cg) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(stage + r * SP + cg + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
    // phase B (coalesced): consecutive threads own consecutive 8-column groups of a row -> full-line global loads/stores
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int idx = threadIdx.x + it * NT;
      const int r = idx / GPR, cg = (idx % GPR) * 8;
      const int row = m0 + r;
      if (row >= p.M || !finish) continue;
      const int nb = n0 + cg;
      float o[8];
      {
        const float4 a = *reinterpret_cast<const float4*>(stage + r * SP + cg);
        const float4 b = *reinterpret_cast<const float4*>(stage + r * SP + cg + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
      }
      if (p.epi == EPI_RES || p.epi == EPI_RES_GATE) {
        float4 r0, r1, g0, g1;
        if (PRE && prefetched) {
          r0 = pre_r[(PRE ? it : 0) * 2]; r1 = pre_r[(PRE ? it : 0) * 2 + 1];
          g0 = pre_g[(PRE ? it : 0) * 2]; g1 = pre_g[(PRE ? it : 0) * 2 + 1];
        } else {
          const float* rr = p.R + (size_t)row * p.ldr + nb;
          r0 = *reinterpret_cast<const float4*>(rr); r1 = *reinterpret_cast<const float4*>(rr + 4);
          if (p.epi == EPI_RES_GATE) {
            const float* gp = p.gate + (size_t)(row / p.rows_per_group) * p.gate_stride + nb;
            g0 = *reinterpret_cast<const float4*>(gp); g1 = *reinterpret_cast<const float4*>(gp + 4);
          }
        }
        if (p.epi == EPI_RES_GATE) {
          o[0] = r0.x + g0.x * o[0]; o[1] = r0.y + g0.y * o[1]; o[2] = r0.z + g0.z * o[2]; o[3] = r0.w + g0.w * o[3];
          o[4] = r1.x + g1.x * o[4]; o[5] = r1.y + g1.y * o[5]; o[6] = r1.z + g1.z * o[6]; o[7] = r1.w + g1.w * o[7];
        } else {
          o[0] += r0.x; o[1] += r0.y; o[2] += r0.z; o[3] += r0.w; o[4] += r1.x; o[5] += r1.y; o[6] += r1.z; o[7] += r1.w;
        }
      }
      if (EXT && p.epi == EPI_GELUBWD16) {     // activation backward fused into the dgrad GEMM: o = dL/dg * GELU'(h), h = p.R
        const float* hr = p.R + (size_t)row * p.ldr + nb;
        const float4 h0 = *reinterpret_cast<const float4*>(hr), h1 = *reinterpret_cast<const float4*>(hr + 4);
        o[0] *= gelu_erf_grad(h0.x); o[1] *= gelu_erf_grad(h0.y); o[2] *= gelu_erf_grad(h0.z); o[3] *= gelu_erf_grad(h0.w);
        o[4] *= gelu_erf_grad(h1.x); o[5] *= gelu_erf_grad(h1.y); o[6] *= gelu_erf_grad(h1.z); o[7] *= gelu_erf_grad(h1.w);
        if (p.colpart) {
          *reinterpret_cast<float4*>(stage + r * SP + cg) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(stage + r * SP + cg + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
      const int orow = p.gi ? (row / p.gi) * p.go + p.goff + row % p.gi : row;
      if (p.C) {
        float* cr = p.C + (size_t)orow * p.ldc + nb;
        *reinterpret_cast<float4*>(cr) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(cr + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (EXT && p.epi == EPI_GELU16) {
#pragma unroll
        for (int t = 0; t < 8; ++t) o[t] = gelu_erf_fast(o[t]);
      }
      if (p.C16) {
        __align__(16) __nv_bfloat16 hi[8];
        __align__(16) __nv_bfloat16 lo[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) split_bf16(o[t], hi[t], lo[t]);
        __nv_bfloat16* ch = p.C16 + (size_t)orow * p.ldc16 + nb;
        *reinterpret_cast<uint4*>(ch) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(ch + p.lo_off) = *reinterpret_cast<const uint4*>(lo);
      }
    }
    if (EXT && p.epi == EPI_GELUBWD16 && p.colpart && finish) {
      // column sums of the emitted operand over this tile's rows (rows past M hold zeros: their A rows were zero-filled by TMA)
      __syncthreads();
      for (int c = threadIdx.x; c < BN; c += NT) {
        float s = 0.f;
#pragma unroll 8
        for (int r = 0; r < BM; ++r) s += stage[r * SP + c];
        p.colpart[(size_t)blockIdx.y * p.N + n0 + c] = s;
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) TC_STAMP(7);
  }
  __syncthreads();
  ktrace(KT_GEMM, 1);
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host side

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
struct TmaEncoder {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeFn fn = nullptr;
  std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> cache;
  char msg[160];

  const char* init() {
    if (fn) return nullptr;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) {
      snprintf(msg, sizeof(msg), "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
      return msg;
    }
    fn = reinterpret_cast<EncodeFn>(f);
    return nullptr;
  }

  // 2-D bf16 row-major [rows, cols] (leading dimension ld elements), box = 64 columns x box_rows, SWIZZLE_128B
  const char* get(const __nv_bfloat16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
    auto key = std::make_tuple((const void*)ptr, rows, cols, ld, box_rows);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return nullptr; }
    if (!fn) return "TMA encoder not initialised";
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (CUresult %d) rows=%d cols=%d ld=%d", (int)r, rows, cols, ld);
      return msg;
    }
    cache.emplace(key, m);
    *out = m;
    return nullptr;
  }
};

template <int BN, int PASSES, bool EXT = false>
inline const char* configure_one() {
  cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN, PASSES, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayout<BN, PASSES>::TOTAL);
  return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}
inline const char* configure_kernels() {
  const char* e;
  if ((e = configure_one<192, 3>())) return e;
  if ((e = configure_one<192, 1>())) return e;
  if ((e = configure_one<128, 3>())) return e;
  if ((e = configure_one<64, 3>())) return e;
  if ((e = configure_one<128, 1>())) return e;
  if ((e = configure_one<64, 1>())) return e;
  if ((e = configure_one<192, 3, true>())) return e;
  if ((e = configure_one<128, 3, true>())) return e;
  if ((e = configure_one<64, 3, true>())) return e;
  return nullptr;
}

template <int BN, int PASSES, bool EXT = false>
inline void launch_one(const CUtensorMap& a, const CUtensorMap& w, const TcParams& p, cudaStream_t st) {
  dim3 grid(p.N / BN, (p.M + BM - 1) / BM, EXT && p.splits > 1 ? p.splits : 1);
  launch_pdl(tc_gemm_kernel<BN, PASSES, EXT>, grid, dim3(THREADS), SmemLayout<BN, PASSES>::TOTAL, st, a, w, p);
}

// returns nullptr on success, an error message otherwise
inline const char* launch_tc_gemm(TmaEncoder& enc, const TcGemm& g, cudaStream_t st) {
  if (g.N % 64 != 0 || g.M < 1 || g.K < 1) return "unsupported GEMM shape (need N % 64 == 0)";
  if ((!g.a_mn || !g.w_mn) && g.K % BK != 0) return "unsupported GEMM shape (a K-major operand needs K % 64 == 0)";
  if ((g.a_mn || g.w_mn) && (g.wg_rows || g.passes != 3)) return "MN-major operands: no weight groups, bf16x3 only";
  if (g.splits > 1 && (!g.sk_ws || !g.sk_cnt || g.splits > (g.K + BK - 1) / BK || g.C16 || g.gi)) return "bad split-K configuration";
  if (g.epi != EPI_NONE && g.epi != EPI_GELU && g.epi != EPI_RES && g.epi != EPI_RES_GATE && g.epi != EPI_GELU16 && g.epi != EPI_GELUBWD16) return "unsupported epilogue";
  // tile-N choice (measured, profiles/r01_experiments.md): 128 columns for the wide GEMMs (N >= 1024), 64 for N = d; the
  // 192-column tile turns the fused QKV (N = 1152) into a single wave when a full batch is launched at once (mt >= 16).
  // A "widest tile that still fills ~1 wave, else narrowest" policy was 11 % slower end to end and no better for training.
  const int mt = (g.M + BM - 1) / BM;
  const bool wide = mt >= 16 && g.N % 192 == 0 && (g.N / 192) * mt <= 148 && (g.N % 128 != 0 || (g.N / 128) * mt > 148);
  // d-wide outputs (N < 1024): 128 columns once there are >= 4 row tiles (A/B at B = 256, three alternations on one box: 2448 vs 2408
  // steps/s: fewer, fatter CTAs leave more SMs to the other chains), 64 columns for the few-row launches where the CTA count is the limit
  int bn = wide ? 192 : (g.N % 128 == 0 && (g.N >= 1024 || mt >= 4)) ? 128 : 64;
  {   // experiment overrides: MDTB200_BN_D (GEMMs with N < 1024), MDTB200_BN_WIDE (N >= 1024)
    static const int ov_d = getenv("MDTB200_BN_D") ? atoi(getenv("MDTB200_BN_D")) : 0;
    static const int ov_w = getenv("MDTB200_BN_WIDE") ? atoi(getenv("MDTB200_BN_WIDE")) : 0;
    const int ov = g.N >= 1024 ? ov_w : ov_d;
    if (g.bn_hint && g.N % g.bn_hint == 0) bn = g.bn_hint;
    if ((ov == 64 || ov == 128 || ov == 192) && g.N % ov == 0) bn = ov;
  }
  CUtensorMap ta, tw;
  const char* e;
  if (g.a_mn) e = enc.get(g.A16, g.K, g.lda16, g.lda16, BK, &ta);
  else e = enc.get(g.A16, g.M, 2 * g.K, g.lda16, BM, &ta);
  if (e) return e;
  if (g.w_mn) e = enc.get(g.W16, g.K, g.ldw16, g.ldw16, BK, &tw);
  else e = enc.get(g.W16, g.w_rows > 0 ? g.w_rows : g.N, 2 * g.K, g.ldw16 > 0 ? g.ldw16 : 2 * g.K, bn, &tw);
  if (e) return e;
  TcParams p{g.bias, g.C, g.ldc, g.C16, g.ldc16, g.lo_off, g.R, g.ldr, g.gate, g.gate_stride, g.rows_per_group > 0 ? g.rows_per_group : 1,
             g.M, g.N, g.K, g.epi, g.trace, g.gi, g.go, g.goff, g.wg_rows, g.wg_stride, g.w_dynamic ? 0 : 1,
             g.a_mn, g.w_mn, g.lda16 / 2, g.ldw16 / 2, g.splits > 1 ? g.splits : 1, g.sk_ws, g.sk_cnt, g.colpart};
  const bool ext = g.a_mn || g.w_mn || g.splits > 1 || g.epi == EPI_GELU16 || g.epi == EPI_GELUBWD16 || g.K % BK != 0;
  if (g.epi == EPI_GELUBWD16 && (!g.R || !g.C16 || g.splits > 1)) return "GELUBWD16 needs R (pre-activation), C16, no split-K";
  if (ext) {
    if (bn == 192) launch_one<192, 3, true>(ta, tw, p, st); else if (bn == 128) launch_one<128, 3, true>(ta, tw, p, st); else launch_one<64, 3, true>(ta, tw, p, st);
  } else if (g.passes == 3) {
    if (bn == 192) launch_one<192, 3>(ta, tw, p, st); else if (bn == 128) launch_one<128, 3>(ta, tw, p, st); else launch_one<64, 3>(ta, tw, p, st);
  } else {
    if (bn == 192) launch_one<192, 1>(ta, tw, p, st); else if (bn == 128) launch_one<128, 1>(ta, tw, p, st); else launch_one<64, 1>(ta, tw, p, st);
  }
  return nullptr;
}

}}  // namespace mdt::tc
