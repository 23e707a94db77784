// Programmatic dependent launch inside a CUDA graph: per-boundary gain for chains of small and of SM-filling kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/exp_pdl scripts/experiments/exp_pdl.cu && gpurun_out/exp_pdl
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <bool PDL>
__global__ void __launch_bounds__(256) work_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int iters) {
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = in[i];
    for (int k = 0; k < iters; ++k) v = fmaf(v, 1.0001f, 0.5f);
    out[i] = v;
  }
}

template <bool PDL>
static void launch(const float* in, float* out, long long n, int iters, int grid, cudaStream_t s) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = PDL ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, work_kernel<PDL>, in, out, n, iters);
}

template <bool PDL>
static float run(float* a, float* b, long long n, int iters, int grid, int chain) {
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < chain; ++i) launch<PDL>(i % 2 ? b : a, i % 2 ? a : b, n, iters, grid, s);
  if (cudaStreamEndCapture(s, &g) != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError())); return -1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, s);
  cudaEventRecord(e0, s);
  for (int w = 0; w < 10; ++w) cudaGraphLaunch(ge, s);
  cudaEventRecord(e1, s);
  cudaStreamSynchronize(s);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 10 / chain * 1000.f;   // us per kernel
}

int main() {
  float *a, *b;
  const long long nmax = 64ll << 20;
  cudaMalloc(&a, nmax * 4); cudaMalloc(&b, nmax * 4);
  cudaMemset(a, 0, nmax * 4); cudaMemset(b, 0, nmax * 4);
  struct Case { const char* name; long long n; int iters; int grid; } cases[] = {
      {"tiny (1 CTA)", 256, 8, 1}, {"small (64 CTAs, 64K elems)", 65536, 8, 64}, {"one wave (592 CTAs, 4M elems)", 4 << 20, 8, 592},
      {"3.5 waves (2048 CTAs, 16M elems)", 16 << 20, 8, 2048}, {"persistent 148 CTAs, 16M elems", 16 << 20, 8, 148}};
  for (auto& c : cases) {
    const float t0 = run<false>(a, b, c.n, c.iters, c.grid, 200);
    const float t1 = run<true>(a, b, c.n, c.iters, c.grid, 200);
    printf("%-36s  plain %7.2f us/kernel   PDL %7.2f us/kernel   gain %5.2f us\n", c.name, t0, t1, t0 - t1);
  }
  return 0;
}
