// Micro-probe (B200): cost of a dependent kernel boundary inside a CUDA graph (plain / PDL) vs a software grid barrier
// inside one persistent kernel.  Decides whether a persistent "step" kernel beats the 46-kernel graph.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency_probe tools/latency_probe.cu && ./latency_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k_empty(float* p) { if (p && threadIdx.x == 9999) p[0] = 1.f; }
__global__ void k_pdl(float* p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 9999) p[0] = 1.f;
}
// touches memory so that consecutive kernels carry a real dependency (read-modify-write of 2.5 MB)
__global__ void k_rmw(float* p, int n) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = p[i] * 1.0001f + 1.f;
}

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < epoch);
  }
  __syncthreads();
}
__global__ void k_persistent(float* p, int n, unsigned* ctr, int iters, int work) {
  unsigned epoch = 0;
  for (int it = 0; it < iters; ++it) {
    if (work) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = p[i] * 1.0001f + 1.f;
    grid_barrier(ctr, epoch);
  }
}

static float run_graph(cudaGraphExec_t g, cudaStream_t s, int reps) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaGraphLaunch(g, s); cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(g, s);
  cudaEventRecord(b, s); cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

template <typename F> static cudaGraphExec_t capture(cudaStream_t s, F f) {
  cudaGraph_t g; cudaGraphExec_t e;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal); f(); cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&e, g, 0); cudaGraphDestroy(g); return e;
}

int main() {
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  const int n = 640 * 1024, N = 500;
  float* p; cudaMalloc(&p, n * 4); cudaMemset(p, 0, n * 4);
  unsigned* ctr; cudaMalloc(&ctr, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  auto launch = [&](auto kern, dim3 g, dim3 b, bool pdl, auto... args) {
    cudaLaunchConfig_t c{}; c.gridDim = g; c.blockDim = b; c.stream = s;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    c.attrs = at; c.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&c, kern, args...);
  };
  struct { const char* name; cudaGraphExec_t g; } tests[] = {
    {"500 x empty kernel (1 CTA)", capture(s, [&] { for (int i = 0; i < N; ++i) launch(k_empty, dim3(1), dim3(32), false, p); })},
    {"500 x empty kernel (148 CTA x 256)", capture(s, [&] { for (int i = 0; i < N; ++i) launch(k_empty, dim3(sms), dim3(256), false, p); })},
    {"500 x pdl kernel (148 CTA x 256), PDL attr", capture(s, [&] { for (int i = 0; i < N; ++i) launch(k_pdl, dim3(sms), dim3(256), true, p); })},
    {"500 x rmw 2.5MB kernel (296 CTA), no PDL attr", capture(s, [&] { for (int i = 0; i < N; ++i) launch(k_rmw, dim3(2 * sms), dim3(256), false, p, n); })},
    {"500 x rmw 2.5MB kernel (296 CTA), PDL attr", capture(s, [&] { for (int i = 0; i < N; ++i) launch(k_rmw, dim3(2 * sms), dim3(256), true, p, n); })},
  };
  for (auto& t : tests) printf("%-52s %8.2f us per node\n", t.name, run_graph(t.g, s, 10) * 1000.f / N);
  for (int work = 0; work < 2; ++work) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaMemsetAsync(ctr, 0, 4, s);
    k_persistent<<<sms, 256, 0, s>>>(p, n, ctr, 10, work); cudaStreamSynchronize(s);
    cudaMemsetAsync(ctr, 0, 4, s);
    cudaEventRecord(a, s);
    k_persistent<<<sms, 256, 0, s>>>(p, n, ctr, N, work);
    cudaEventRecord(b, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("persistent kernel, 500 grid barriers (%s)         %8.2f us per phase   [%s]\n", work ? "rmw 2.5MB each" : "no work", ms * 1000.f / N, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
