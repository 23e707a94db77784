// HBM-bound kernels of the training step: optimiser / EMA updates over flat fp32 buffers, the
// crop + bilinear resize of AsterInferer.convert_inputs and small fused element-wise passes.
// All are grid-stride, 128-bit vectorised where alignment allows, and sized to a multiple of the
// SM count.
#include <initializer_list>

#include "common.cuh"
#include "host_util.h"

namespace tbg {

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = static_cast<long long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ----------------------------------------------------------------------------------------------
// tf.keras.optimizers.Adam (optimizer_v2, non-amsgrad) on a flat buffer:
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr_t * m / (sqrt(v) + eps)
// with lr_t = lr*sqrt(1-b2^t)/(1-b1^t) computed on the host (train.py:58-75).
// ----------------------------------------------------------------------------------------------
// Pointers may start at any 4-byte boundary as long as they share the same 16-byte phase: the
// first `head` and the last few elements are handled with scalar accesses, the body with float4.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, int head, float lr_t, const float* __restrict__ lr_t_dev,
                            float b1, float b2, float eps) {
  if (lr_t_dev != nullptr) lr_t = __ldg(lr_t_dev);
  const long long n4 = (n - head) >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float4* p4 = reinterpret_cast<float4*>(p + head);
  const float4* g4 = reinterpret_cast<const float4*>(g + head);
  float4* m4 = reinterpret_cast<float4*>(m + head);
  float4* v4 = reinterpret_cast<float4*>(v + head);
  for (long long i = tid; i < n4; i += stride) {
    float4 pp = p4[i];
    const float4 gg = g4[i];
    float4 mm = m4[i];
    float4 vv = v4[i];
#define TBG_ADAM1(c)                               \
  mm.c = b1 * mm.c + (1.f - b1) * gg.c;            \
  vv.c = b2 * vv.c + (1.f - b2) * gg.c * gg.c;     \
  pp.c -= lr_t * mm.c / (sqrtf(vv.c) + eps);
    TBG_ADAM1(x) TBG_ADAM1(y) TBG_ADAM1(z) TBG_ADAM1(w)
#undef TBG_ADAM1
    p4[i] = pp;
    m4[i] = mm;
    v4[i] = vv;
  }
  // scalar head [0, head) and tail [head + 4*n4, n)
  const long long tail0 = head + (n4 << 2);
  const long long n_scalar = head + (n - tail0);
  for (long long s = tid; s < n_scalar; s += stride) {
    const long long i = s < head ? s : tail0 + (s - head);
    const float gg = g[i];
    const float mm = b1 * m[i] + (1.f - b1) * gg;
    const float vv = b2 * v[i] + (1.f - b2) * gg * gg;
    m[i] = mm;
    v[i] = vv;
    p[i] -= lr_t * mm / (sqrtf(vv) + eps);
  }
}

// dst = src + (dst - src) * beta     (generator.py:48-59, lerp(sw, cw, beta))
__global__ void ema_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, int head, float beta) {
  const long long n4 = (n - head) >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  float4* d4 = reinterpret_cast<float4*>(dst + head);
  const float4* s4 = reinterpret_cast<const float4*>(src + head);
  for (long long i = tid; i < n4; i += stride) {
    float4 d = d4[i];
    const float4 s = s4[i];
    d.x = s.x + (d.x - s.x) * beta;
    d.y = s.y + (d.y - s.y) * beta;
    d.z = s.z + (d.z - s.z) * beta;
    d.w = s.w + (d.w - s.w) * beta;
    d4[i] = d;
  }
  const long long tail0 = head + (n4 << 2);
  const long long n_scalar = head + (n - tail0);
  for (long long s = tid; s < n_scalar; s += stride) {
    const long long i = s < head ? s : tail0 + (s - head);
    dst[i] = src[i] + (dst[i] - src[i]) * beta;
  }
}

// Common scalar head so that (ptr + head) is 16-byte aligned for every pointer, or n (all scalar)
// when the pointers do not share a 16-byte phase.
static long long common_head(long long n, std::initializer_list<const void*> ptrs) {
  uintptr_t phase = reinterpret_cast<uintptr_t>(*ptrs.begin()) & 15;
  for (const void* q : ptrs)
    if ((reinterpret_cast<uintptr_t>(q) & 15) != phase) return n;
  long long head = ((16 - static_cast<long long>(phase)) & 15) >> 2;
  return head < n ? head : n;
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr_t,
                             const float* lr_t_dev, float beta1, float beta2, float eps, void* stream_v) {
  TBG_CHECK_ARG(p && g && m && v && n >= 0, "tbg_adam_step: bad arguments");
  TBG_CHECK_ARG(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) & 3) == 0,
                "tbg_adam_step: buffers must be 4-byte aligned");
  if (n == 0) return TBG_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int head = static_cast<int>(common_head(n, {p, g, m, v}));
  adam_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, stream>>>(p, g, m, v, n, head, lr_t, lr_t_dev, beta1, beta2, eps);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_ema_step(float* dst, const float* src, long long n, float beta, void* stream_v) {
  TBG_CHECK_ARG(dst && src && n >= 0, "tbg_ema_step: bad arguments");
  TBG_CHECK_ARG(((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 3) == 0,
                "tbg_ema_step: buffers must be 4-byte aligned");
  if (n == 0) return TBG_OK;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int head = static_cast<int>(common_head(n, {dst, src}));
  ema_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, stream>>>(dst, src, n, head, beta);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

// ---------------------------------------------------------------------------------------------
// AsterInferer.convert_inputs (aster_inferer.py:153-190) as one gather kernel forward and one scatter kernel
// backward: NCHW fp32 image -> crop at first_blank * char_width -> bilinear resize (tf.image.resize: half-pixel
// centres, no antialias) -> NHWC fp32 [B, oh, ow, 3].
//   src = (o + 0.5) * in / out - 0.5 ; i0 = clamp(floor(src)), i1 = clamp(ceil(src)), weights (1 - frac, frac)
// The crop width of sample b is floor(first_blank(b) * cw_num / cw_den) clamped to [1, W] (W when no blank label).
// ---------------------------------------------------------------------------------------------
namespace tbg {

__device__ __forceinline__ int crop_width_of(const int* __restrict__ labels, int b, int mcn, int blank, int cw_num,
                                             int cw_den, int W) {
  int first = -1;
  for (int i = 0; i < mcn; ++i)
    if (__ldg(labels + static_cast<size_t>(b) * mcn + i) == blank) {
      first = i;
      break;
    }
  if (first < 0) return W;
  long long wc = (static_cast<long long>(first) * cw_num) / cw_den;
  if (wc < 1) wc = 1;
  if (wc > W) wc = W;
  return static_cast<int>(wc);
}

__device__ __forceinline__ void resize_taps(int o, int in_size, int out_size, int& i0, int& i1, float& w0, float& w1) {
  const float src = (o + 0.5f) * (static_cast<float>(in_size) / static_cast<float>(out_size)) - 0.5f;
  const float f0 = floorf(src);
  const float lerp = src - f0;
  const int hi = in_size - 1;
  i0 = min(max(static_cast<int>(f0), 0), hi);
  i1 = min(max(static_cast<int>(ceilf(src)), 0), hi);
  w0 = 1.f - lerp;
  w1 = lerp;
}

__global__ void __launch_bounds__(256)
crop_resize_fwd_kernel(const float* __restrict__ img, const int* __restrict__ labels, float* __restrict__ out, int B, int H,
                       int W, int oh, int ow, int mcn, int blank, int cw_num, int cw_den) {
  const long long n = static_cast<long long>(B) * oh * ow;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(e % ow), y = static_cast<int>((e / ow) % oh), b = static_cast<int>(e / (static_cast<long long>(ow) * oh));
    const int wc = crop_width_of(labels, b, mcn, blank, cw_num, cw_den, W);
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    resize_taps(y, H, oh, y0, y1, wy0, wy1);
    resize_taps(x, wc, ow, x0, x1, wx0, wx1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* pl = img + (static_cast<size_t>(b) * 3 + c) * H * W;
      const float v = wy0 * (wx0 * __ldg(pl + y0 * W + x0) + wx1 * __ldg(pl + y0 * W + x1)) +
                      wy1 * (wx0 * __ldg(pl + y1 * W + x0) + wx1 * __ldg(pl + y1 * W + x1));
      out[e * 3 + c] = v;
    }
  }
}

__global__ void __launch_bounds__(256)
crop_resize_bwd_kernel(const float* __restrict__ g, const int* __restrict__ labels, float* __restrict__ gimg, int B, int H,
                       int W, int oh, int ow, int mcn, int blank, int cw_num, int cw_den) {
  const long long n = static_cast<long long>(B) * oh * ow;
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(e % ow), y = static_cast<int>((e / ow) % oh), b = static_cast<int>(e / (static_cast<long long>(ow) * oh));
    const int wc = crop_width_of(labels, b, mcn, blank, cw_num, cw_den, W);
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    resize_taps(y, H, oh, y0, y1, wy0, wy1);
    resize_taps(x, wc, ow, x0, x1, wx0, wx1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gv = __ldg(g + e * 3 + c);
      float* pl = gimg + (static_cast<size_t>(b) * 3 + c) * H * W;
      atomicAdd(pl + y0 * W + x0, gv * wy0 * wx0);
      atomicAdd(pl + y0 * W + x1, gv * wy0 * wx1);
      atomicAdd(pl + y1 * W + x0, gv * wy1 * wx0);
      atomicAdd(pl + y1 * W + x1, gv * wy1 * wx1);
    }
  }
}

}  // namespace tbg

extern "C" int tbg_crop_resize_fwd(const float* img, const int* labels, float* out, int B, int H, int W, int oh, int ow,
                                   int mcn, int blank, int cw_num, int cw_den, void* stream_v) {
  TBG_CHECK_ARG(img && labels && out, "tbg_crop_resize_fwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && oh >= 1 && ow >= 1 && mcn >= 1 && cw_num >= 1 && cw_den >= 1,
                "tbg_crop_resize_fwd: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n = static_cast<long long>(B) * oh * ow;
  tbg::crop_resize_fwd_kernel<<<static_cast<int>((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, stream>>>(
      img, labels, out, B, H, W, oh, ow, mcn, blank, cw_num, cw_den);
  tbg::count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_crop_resize_bwd(const float* g, const int* labels, float* gimg, int B, int H, int W, int oh, int ow,
                                   int mcn, int blank, int cw_num, int cw_den, void* stream_v) {
  TBG_CHECK_ARG(g && labels && gimg, "tbg_crop_resize_bwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && oh >= 1 && ow >= 1 && mcn >= 1 && cw_num >= 1 && cw_den >= 1,
                "tbg_crop_resize_bwd: bad shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n = static_cast<long long>(B) * oh * ow;
  tbg::crop_resize_bwd_kernel<<<static_cast<int>((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, stream>>>(
      g, labels, gimg, B, H, W, oh, ow, mcn, blank, cw_num, cw_den);
  tbg::count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

// ---------------------------------------------------------------------------------------------
// Batched loader transform (dataset_utils/training_data_loader.py:64-86) on the device.  For sample b the BGR uint8 image
// src + offsets[b] (src_h[b] x src_w[b] x 3, HWC) is resized bilinearly to dst_w[b] x H (cv2.resize INTER_LINEAR: pixel
// centres, edge clamp, no antialias; the exact 2:1 decimation uses the 2 x 2 mean as cv2 does), rounded to the uint8 grid,
// scaled to [-1, 1] (v / 127.5 - 1), zero-padded on the right up to W (cv2.copyMakeBorder) and written CHW: out fp32
// [B,3,H,W].  One thread per output pixel; the host ships all images of a batch in ONE pinned buffer.
// ---------------------------------------------------------------------------------------------
namespace tbg {

__global__ void __launch_bounds__(256)
batch_resize_normalize_kernel(const unsigned char* __restrict__ src, const long long* __restrict__ offsets,
                              const int* __restrict__ src_h, const int* __restrict__ src_w, const int* __restrict__ dst_w,
                              float* __restrict__ out, int B, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % W);
    const int y = static_cast<int>((i / W) % H);
    const int b = static_cast<int>(i / (static_cast<long long>(W) * H));
    float v[3] = {0.f, 0.f, 0.f};
    const int dw = dst_w[b];
    if (x < dw) {
      const int sh = src_h[b], sw = src_w[b];
      const unsigned char* img = src + offsets[b];
      float r[3];
      if (sw == 2 * dw && sh == 2 * H) {
        for (int c = 0; c < 3; ++c) {
          const int s = img[((2 * y) * sw + 2 * x) * 3 + c] + img[((2 * y) * sw + 2 * x + 1) * 3 + c] +
                        img[((2 * y + 1) * sw + 2 * x) * 3 + c] + img[((2 * y + 1) * sw + 2 * x + 1) * 3 + c];
          r[c] = static_cast<float>((s + 2) >> 2);
        }
      } else {
        float fx = (x + 0.5f) * (static_cast<float>(sw) / dw) - 0.5f;
        float fy = (y + 0.5f) * (static_cast<float>(sh) / H) - 0.5f;
        int x0 = static_cast<int>(floorf(fx)), y0 = static_cast<int>(floorf(fy));
        fx -= x0;
        fy -= y0;
        if (x0 < 0) { x0 = 0; fx = 0.f; }
        if (x0 >= sw - 1) { x0 = sw - 1; fx = 0.f; }
        if (y0 < 0) { y0 = 0; fy = 0.f; }
        if (y0 >= sh - 1) { y0 = sh - 1; fy = 0.f; }
        const int x1 = min(x0 + 1, sw - 1), y1 = min(y0 + 1, sh - 1);
        for (int c = 0; c < 3; ++c) {
          const float p00 = img[(y0 * sw + x0) * 3 + c], p01 = img[(y0 * sw + x1) * 3 + c];
          const float p10 = img[(y1 * sw + x0) * 3 + c], p11 = img[(y1 * sw + x1) * 3 + c];
          const float top = p00 + (p01 - p00) * fx, bot = p10 + (p11 - p10) * fx;
          r[c] = rintf(top + (bot - top) * fy);
        }
      }
      for (int c = 0; c < 3; ++c) v[c] = r[c] / 127.5f - 1.f;
    }
    const size_t plane = static_cast<size_t>(H) * W;
    float* o = out + static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(y) * W + x;
    o[0] = v[0];
    o[plane] = v[1];
    o[2 * plane] = v[2];
  }
}

}  // namespace tbg

extern "C" int tbg_batch_resize_normalize(const unsigned char* src, const long long* offsets, const int* src_h,
                                          const int* src_w, const int* dst_w, float* out, int B, int H, int W,
                                          void* stream_v) {
  TBG_CHECK_ARG(src && offsets && src_h && src_w && dst_w && out, "tbg_batch_resize_normalize: null pointer");
  TBG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "tbg_batch_resize_normalize: bad shape B=%d H=%d W=%d", B, H, W);
  const long long total = static_cast<long long>(B) * H * W;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  tbg::batch_resize_normalize_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(src, offsets, src_h, src_w,
                                                                                                dst_w, out, B, H, W);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
