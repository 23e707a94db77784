// Batch-statistic kernels of the discriminator head and the R1 regulariser.
//   * MinibatchStd.call (mini_batch_std.py:10-35): groups of G = min(4, B) samples {m, m + B/G, ...}; per feature the
//     biased standard deviation over the group (sqrt(var + 1e-8)), averaged over all features (C, H, W) -> one value per
//     group, appended to every sample of the group as one extra constant channel.
//   * _r1_reg (training_step.py:363-372): per-sample squared L2 norm of the image gradient.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.f;
  if (warp == 0) {
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) scratch[0] = t;
  }
  __syncthreads();
  return scratch[0];
}

// One CTA per (call, group-slot m).  x: bf16 [n_calls * B, F] (F = H*W*C features of one sample, NHWC order),
// xcat: bf16 [n_calls * B, HW, Cpad] = [x | std | zeros], stat: fp32 [n_calls * B] (the statistic of each sample's group).
__global__ void __launch_bounds__(256)
minibatch_std_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ xcat, float* __restrict__ stat,
                         int B, int G, int HW, int C, int Cpad) {
  __shared__ float scratch[32];
  const int groups = B / G;                      // samples of a group: m, m + groups, ...
  const int call = blockIdx.x / groups, m = blockIdx.x - call * groups;
  const size_t F = static_cast<size_t>(HW) * C;
  const __nv_bfloat16* xb = x + static_cast<size_t>(call) * B * F;
  float acc = 0.f;
  for (size_t f = threadIdx.x; f < F; f += blockDim.x) {
    float v[4], mean = 0.f;
    for (int g = 0; g < G; ++g) {
      v[g] = __bfloat162float(xb[(static_cast<size_t>(g) * groups + m) * F + f]);
      mean += v[g];
    }
    mean /= G;
    float var = 0.f;
    for (int g = 0; g < G; ++g) var = fmaf(v[g] - mean, v[g] - mean, var);
    acc += sqrtf(var / G + 1e-8f);
  }
  const float sd = block_sum(acc, scratch) / static_cast<float>(F);
  const __nv_bfloat16 sd_b = __float2bfloat16(sd);
  for (int g = 0; g < G; ++g) {
    const size_t n = static_cast<size_t>(call) * B + static_cast<size_t>(g) * groups + m;
    if (threadIdx.x == 0) stat[n] = sd;
    const __nv_bfloat16* src = x + n * F;
    __nv_bfloat16* dst = xcat + n * HW * Cpad;
    for (int e = threadIdx.x; e < HW * Cpad; e += blockDim.x) {
      const int p = e / Cpad, c = e - p * Cpad;
      dst[e] = c < C ? src[static_cast<size_t>(p) * C + c] : (c == C ? sd_b : __float2bfloat16(0.f));
    }
  }
}

// gx[n, p, c] = gxcat[n, p, c] + gstat(group of n) * (x[n,f] - mean_f) / (G * sd_f * F),
// gstat(group) = sum over the group's samples and pixels of gxcat[n, p, C] (the gradient of the appended channel).
__global__ void __launch_bounds__(256)
minibatch_std_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gxcat,
                         __nv_bfloat16* __restrict__ gx, int B, int G, int HW, int C, int Cpad) {
  __shared__ float scratch[32];
  const int groups = B / G;
  const int call = blockIdx.x / groups, m = blockIdx.x - call * groups;
  const size_t F = static_cast<size_t>(HW) * C;
  float part = 0.f;
  for (int e = threadIdx.x; e < G * HW; e += blockDim.x) {
    const int g = e / HW, p = e - g * HW;
    const size_t n = static_cast<size_t>(call) * B + static_cast<size_t>(g) * groups + m;
    part += __bfloat162float(gxcat[(n * HW + p) * Cpad + C]);
  }
  const float gstat = block_sum(part, scratch);
  const float k = gstat / (static_cast<float>(G) * static_cast<float>(F));
  for (size_t f = threadIdx.x; f < F; f += blockDim.x) {
    const int p = static_cast<int>(f / C), c = static_cast<int>(f - static_cast<size_t>(p) * C);
    float v[4], mean = 0.f;
    size_t n[4];
    for (int g = 0; g < G; ++g) {
      n[g] = static_cast<size_t>(call) * B + static_cast<size_t>(g) * groups + m;
      v[g] = __bfloat162float(x[n[g] * F + f]);
      mean += v[g];
    }
    mean /= G;
    float var = 0.f;
    for (int g = 0; g < G; ++g) var = fmaf(v[g] - mean, v[g] - mean, var);
    const float inv_sd = rsqrtf(var / G + 1e-8f);
    for (int g = 0; g < G; ++g) {
      const float gc = __bfloat162float(gxcat[(n[g] * HW + p) * Cpad + c]);
      gx[n[g] * F + f] = __float2bfloat16(gc + k * (v[g] - mean) * inv_sd);
    }
  }
}

// out[b] = sum over the sample's elements of g^2 (fp32 NCHW gradient image)
__global__ void __launch_bounds__(256)
r1_sqnorm_kernel(const float* __restrict__ g, float* __restrict__ out, long long per_sample) {
  __shared__ float scratch[32];
  const float* gb = g + static_cast<size_t>(blockIdx.x) * per_sample;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < per_sample; i += blockDim.x) acc = fmaf(gb[i], gb[i], acc);
  const float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}
// gg[b, i] = 2 * g[b, i] * gout[b]
__global__ void r1_sqnorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ gout, float* __restrict__ gg,
                                     long long per_sample, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    gg[i] = 2.f * g[i] * gout[i / per_sample];
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_minibatch_std_fwd(const void* x, void* xcat, float* stat, int B, int n_calls, int group_size, int HW,
                                     int C, int Cpad, void* stream_v) {
  TBG_CHECK_ARG(x && xcat && stat, "tbg_minibatch_std_fwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && n_calls >= 1 && HW >= 1 && C >= 1 && Cpad > C, "tbg_minibatch_std_fwd: bad shape (Cpad must exceed C)");
  const int G = group_size < B ? group_size : B;
  TBG_CHECK_ARG(G >= 1 && G <= 4 && B % G == 0, "tbg_minibatch_std_fwd: batch %d must be a multiple of the group size %d (<= 4)", B, G);
  minibatch_std_fwd_kernel<<<n_calls * (B / G), 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(xcat), stat, B, G, HW, C, Cpad);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_minibatch_std_bwd(const void* x, const void* gxcat, void* gx, int B, int n_calls, int group_size, int HW,
                                     int C, int Cpad, void* stream_v) {
  TBG_CHECK_ARG(x && gxcat && gx, "tbg_minibatch_std_bwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && n_calls >= 1 && HW >= 1 && C >= 1 && Cpad > C, "tbg_minibatch_std_bwd: bad shape (Cpad must exceed C)");
  const int G = group_size < B ? group_size : B;
  TBG_CHECK_ARG(G >= 1 && G <= 4 && B % G == 0, "tbg_minibatch_std_bwd: batch %d must be a multiple of the group size %d (<= 4)", B, G);
  minibatch_std_bwd_kernel<<<n_calls * (B / G), 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(gxcat),
      reinterpret_cast<__nv_bfloat16*>(gx), B, G, HW, C, Cpad);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_r1_sqnorm(const float* g, float* out, int B, long long per_sample, void* stream_v) {
  TBG_CHECK_ARG(g && out && B >= 1 && per_sample >= 1, "tbg_r1_sqnorm: bad arguments");
  r1_sqnorm_kernel<<<B, 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(g, out, per_sample);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_r1_sqnorm_bwd(const float* g, const float* gout, float* gg, int B, long long per_sample, void* stream_v) {
  TBG_CHECK_ARG(g && gout && gg && B >= 1 && per_sample >= 1, "tbg_r1_sqnorm_bwd: bad arguments");
  const long long total = per_sample * B;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  r1_sqnorm_bwd_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(g, gout, gg, per_sample, total);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
