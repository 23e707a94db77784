// RGB branch of the synthesis network and the generic resampling op.
//   * tbg_upfirdn2d: the reference's one native op, UpFirDn2D (upfirdn_2d.cu:232-324; semantics upfirdn_2d_v2.py:249-305):
//     zero-insert upsample -> pad / crop -> FIR (correlation with the flipped kernel) -> decimate, on
//     [major, inH, inW, minor] tensors.  Polyphase form: output sample o reads only the taps whose upsampled
//     coordinate lands on an input sample, from an input patch staged in shared memory per 8 x 32 output tile.
//   * tbg_torgb_skip_fwd: ToRGB.call (to_rgb.py:28-33) fused with the skip connection of SynthesisBlock.call
//     (synthesis_block.py:152: y = upsample_2d(y_prev) + to_rgb(x); upsample_2d = zero-insert x2, pad 2/1, FIR
//     outer([1,3,3,1])/64 * 4, upfirdn_2d_v2.py:58-62), and for the last block with mask_text_box (utils/utils.py:11-45)
//     and the NHWC -> NCHW layout change of the image.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct UpfirdnGeom {
  int upx, upy, downx, downy, padx0, pady0;
  int major, inH, inW, minor, kH, kW, outH, outW;
  int tile_rows, tile_cols, mchunk;    // staged input patch (pixels) and minor elements per pass
};

static constexpr int kUfTX = 32, kUfTY = 8;

__device__ __forceinline__ int floordiv(int a, int b) {   // b > 0
  const int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ int posmod(int a, int b) {     // b > 0
  const int r = a % b;
  return r < 0 ? r + b : r;
}

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// y[oy, ox] = sum_{ky, kx} U[oy*downy + ky - pady0, ox*downx + kx - padx0] * k[kH-1-ky, kW-1-kx],
// U[u, v] = x[u/upy, v/upx] when both are exact multiples inside the input, else 0.
template <typename T>
__global__ void __launch_bounds__(kUfTX * kUfTY)
upfirdn2d_tiled_kernel(const T* __restrict__ x, const float* __restrict__ k, T* __restrict__ y, const UpfirdnGeom g) {
  extern __shared__ float smem_f[];
  float* sk = smem_f;                                // flipped taps [kH][kW]
  float* tile = smem_f + g.kH * g.kW;                // [tile_rows][tile_cols][mchunk]
  const int tid = threadIdx.y * kUfTX + threadIdx.x;
  for (int i = tid; i < g.kH * g.kW; i += kUfTX * kUfTY) {
    const int ky = i / g.kW, kx = i - ky * g.kW;
    sk[i] = k[(g.kH - 1 - ky) * g.kW + (g.kW - 1 - kx)];
  }
  const int mj = blockIdx.z;
  const int oy0 = blockIdx.y * kUfTY, ox0 = blockIdx.x * kUfTX;
  const int iy0 = floordiv(oy0 * g.downy - g.pady0, g.upy);        // first input row the tile can touch
  const int ix0 = floordiv(ox0 * g.downx - g.padx0, g.upx);
  const int oy = oy0 + threadIdx.y, ox = ox0 + threadIdx.x;
  const bool live = oy < g.outH && ox < g.outW;
  const int by = oy * g.downy - g.pady0, bx = ox * g.downx - g.padx0;
  const int ky_first = posmod(-by, g.upy), kx_first = posmod(-bx, g.upx);
  for (int m0 = 0; m0 < g.minor; m0 += g.mchunk) {
    const int mc = min(g.mchunk, g.minor - m0);
    __syncthreads();
    for (int i = tid; i < g.tile_rows * g.tile_cols * mc; i += kUfTX * kUfTY) {
      const int m = i % mc, c = (i / mc) % g.tile_cols, r = i / (mc * g.tile_cols);
      const int iy = iy0 + r, ix = ix0 + c;
      float v = 0.f;
      if (iy >= 0 && iy < g.inH && ix >= 0 && ix < g.inW)
        v = ld_f<T>(x + ((static_cast<size_t>(mj) * g.inH + iy) * g.inW + ix) * g.minor + m0 + m);
      tile[(r * g.tile_cols + c) * g.mchunk + m] = v;
    }
    __syncthreads();
    if (!live) continue;
    for (int m = 0; m < mc; ++m) {
      float acc = 0.f;
      for (int ky = ky_first; ky < g.kH; ky += g.upy) {
        const int r = (by + ky) / g.upy - iy0;              // exact division: by + ky is a multiple of upy
        for (int kx = kx_first; kx < g.kW; kx += g.upx) {
          const int c = (bx + kx) / g.upx - ix0;
          acc = fmaf(tile[(r * g.tile_cols + c) * g.mchunk + m], sk[ky * g.kW + kx], acc);
        }
      }
      st_f<T>(y + ((static_cast<size_t>(mj) * g.outH + oy) * g.outW + ox) * g.minor + m0 + m, acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// ToRGB + skip.  x bf16 [B, H, W, C]; ws fp32 [B, C, 3] (style-scaled 1x1 weights, no demodulation); bias fp32 [3];
// y_prev fp32 NHWC [B, H/2, W/2, 3] or null; words int32 [B, mcn] or null (mask_text_box: column w of sample b is kept
// iff words[b, floor(w*mcn/W)] != 0); out fp32 NHWC [B,H,W,3] or NCHW [B,3,H,W].
// One CTA per (sample, run of pixels): the sample's weights sit in registers, LPP = C/8 lanes share a pixel.
// upsample_2d per axis: out[2q] = (x[q-1] + 3 x[q]) / 4, out[2q+1] = (3 x[q] + x[q+1]) / 4 (zero outside).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8_bf16(const uint4& u, float (&f)[8]) {
  // bf16 -> fp32 is a 16-bit shift: one SHL / LOP per value instead of the PRMT pairs of __bfloat1622float2
  f[0] = __uint_as_float(u.x << 16); f[1] = __uint_as_float(u.x & 0xffff0000u);
  f[2] = __uint_as_float(u.y << 16); f[3] = __uint_as_float(u.y & 0xffff0000u);
  f[4] = __uint_as_float(u.z << 16); f[5] = __uint_as_float(u.z & 0xffff0000u);
  f[6] = __uint_as_float(u.w << 16); f[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int LPP, int CV>   // lanes per pixel (<= 32), 8-channel vectors per lane (C = LPP * CV * 8)
__global__ void __launch_bounds__(256)
torgb_skip_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ ws, const float* __restrict__ bias,
                      const float* __restrict__ y_prev, const int* __restrict__ words, float* __restrict__ out, int H, int W,
                      int mcn, int nchw, int pix_per_cta) {
  const int b = blockIdx.y;
  const int hw = H * W;
  const int sub = threadIdx.x % LPP;                 // lane inside the pixel group
  const int grp = threadIdx.x / LPP, ngrp = blockDim.x / LPP;
  constexpr int c8 = LPP * CV;
  float w[CV][8][3];
#pragma unroll
  for (int v = 0; v < CV; ++v) {
    const float* wp = ws + (static_cast<size_t>(b) * c8 + v * LPP + sub) * 24;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) w[v][i][j] = __ldg(wp + i * 3 + j);
  }
  const float b0 = bias ? __ldg(bias) : 0.f, b1 = bias ? __ldg(bias + 1) : 0.f, b2 = bias ? __ldg(bias + 2) : 0.f;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  // epilogue of one pixel (lane 0 of its group): bias, upsampled skip, mask, store
  auto finish = [&](int p, float a0, float a1, float a2) {
    const size_t pix = static_cast<size_t>(b) * hw + p;
    const int py = p / W, px = p - py * W;
    float r0 = a0 + b0, r1 = a1 + b1, r2 = a2 + b2;
    if (y_prev) {
      const int h2 = H >> 1, w2 = W >> 1;
      const int qy = py >> 1, qx = px >> 1;
      // rows / columns and weights of the two contributing low-resolution samples per axis
      const int ya = (py & 1) ? qy : qy - 1, yb = (py & 1) ? qy + 1 : qy;
      const int xa = (px & 1) ? qx : qx - 1, xb = (px & 1) ? qx + 1 : qx;
      const float wya = (py & 1) ? 0.75f : 0.25f, wyb = 1.f - wya;
      const float wxa = (px & 1) ? 0.75f : 0.25f, wxb = 1.f - wxa;
      const float* yp = y_prev + static_cast<size_t>(b) * h2 * w2 * 3;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
      auto tap = [&](int yy, int xx, float wt) {
        if (yy >= 0 && yy < h2 && xx >= 0 && xx < w2) {
          const float* q = yp + (static_cast<size_t>(yy) * w2 + xx) * 3;
          s0 = fmaf(__ldg(q), wt, s0);
          s1 = fmaf(__ldg(q + 1), wt, s1);
          s2 = fmaf(__ldg(q + 2), wt, s2);
        }
      };
      tap(ya, xa, wya * wxa);
      tap(ya, xb, wya * wxb);
      tap(yb, xa, wyb * wxa);
      tap(yb, xb, wyb * wxb);
      r0 += s0; r1 += s1; r2 += s2;
    }
    if (words) {
      const int ch = static_cast<int>((static_cast<long long>(px) * mcn) / W);
      if (__ldg(words + b * mcn + ch) == 0) { r0 = 0.f; r1 = 0.f; r2 = 0.f; }
    }
    if (nchw) {
      float* o = out + static_cast<size_t>(b) * 3 * hw + p;
      o[0] = r0; o[hw] = r1; o[2 * static_cast<size_t>(hw)] = r2;
    } else {
      float* o = out + pix * 3;
      o[0] = r0; o[1] = r1; o[2] = r2;
    }
  };
  // U pixels per group and iteration: U independent 16-byte loads in flight per lane.  The butterfly leaves every lane
  // of the group with the three sums, so lane u finishes pixel u: ONE pass through the (long, divergent) epilogue per
  // iteration with U of LPP lanes active instead of U passes with one lane each.
  constexpr int U = (CV == 1) ? 8 : (CV == 2 ? 4 : 2);
  static_assert(U <= LPP, "one finishing lane per pixel of the iteration");
  for (int pb = p0; pb < p1; pb += ngrp * U) {        // uniform trip count per CTA: the shuffles below need every lane
    uint4 xv[U][CV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * ngrp + grp;
      const size_t pix = static_cast<size_t>(b) * hw + (p < p1 ? p : p0);
#pragma unroll
      for (int v = 0; v < CV; ++v) xv[u][v] = __ldg(x + pix * c8 + v * LPP + sub);
    }
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int v = 0; v < CV; ++v) {
        float f[8];
        unpack8_bf16(xv[u][v], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a0 = fmaf(f[i], w[v][i][0], a0);
          a1 = fmaf(f[i], w[v][i][1], a1);
          a2 = fmaf(f[i], w[v][i][2], a2);
        }
      }
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
      }
      if (sub == u) { m0 = a0; m1 = a1; m2 = a2; }
    }
    const int pm = pb + sub * ngrp + grp;
    if (sub < U && pm < p1) finish(pm, m0, m1, m2);
  }
}

// Gradient of the final image (fp32 NCHW [B,3,H,W]) -> masked NHWC [B,H,W,3]: the adjoint of the layout change and of
// mask_text_box in tbg_torgb_skip_fwd.
__global__ void image_grad_nhwc_kernel(const float* __restrict__ g, const int* __restrict__ words, float* __restrict__ out,
                                       int B, int H, int W, int mcn) {
  const long long total = static_cast<long long>(B) * H * W;
  const int hw = H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / hw), p = static_cast<int>(i - static_cast<long long>(b) * hw);
    const int px = p % W;
    bool keep = true;
    if (words) keep = __ldg(words + b * mcn + static_cast<int>((static_cast<long long>(px) * mcn) / W)) != 0;
    const float* gp = g + static_cast<size_t>(b) * 3 * hw + p;
    float* o = out + i * 3;
    o[0] = keep ? gp[0] : 0.f;
    o[1] = keep ? gp[hw] : 0.f;
    o[2] = keep ? gp[2 * static_cast<size_t>(hw)] : 0.f;
  }
}

// Pixels per CTA of tbg_torgb_skip_fwd: the number of runs per sample (1 .. hw / 64) that minimises
// waves x (iterations per CTA + a fixed per-CTA cost of ~6 iterations for the weight loads and the launch slot).
static int torgb_run(int B, int hw, int slots, int pix_per_iter) {
  int best_run = hw;
  double best = 1e30;
  const int max_runs = hw / 64 > 1 ? hw / 64 : 1;
  for (int r = 1; r <= max_runs; ++r) {
    const int run = (hw + r - 1) / r;
    const long long ctas = static_cast<long long>((hw + run - 1) / run) * B;
    const double waves = static_cast<double>((ctas + slots - 1) / slots);
    const double cost = waves * ((run + pix_per_iter - 1) / pix_per_iter + 6.0);
    if (cost < best) { best = cost; best_run = run; }
    if (ctas > 8LL * slots) break;
  }
  return best_run;
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_upfirdn2d(const void* x, const float* k, void* y, int dtype_bf16, int major, int inH, int inW,
                             int minor, int kH, int kW, int upx, int upy, int downx, int downy, int padx0, int padx1,
                             int pady0, int pady1, void* stream_v) {
  // same argument contract as UpFirDn2DOp::Compute (upfirdn_2d.cu:241-266)
  TBG_CHECK_ARG(x && k && y, "tbg_upfirdn2d: null pointer");
  TBG_CHECK_ARG(upx >= 1 && upy >= 1, "upx and upy must be at least 1x1");
  TBG_CHECK_ARG(downx >= 1 && downy >= 1, "downx and downy must be at least 1x1");
  TBG_CHECK_ARG(kW >= 1 && kH >= 1 && kW * kH <= 1024, "kernel must be between 1x1 and 1024 taps");
  TBG_CHECK_ARG(major >= 1 && inH >= 1 && inW >= 1 && minor >= 1, "input must have rank 4 with positive dims");
  UpfirdnGeom g;
  g.upx = upx; g.upy = upy; g.downx = downx; g.downy = downy; g.padx0 = padx0; g.pady0 = pady0;
  g.major = major; g.inH = inH; g.inW = inW; g.minor = minor; g.kH = kH; g.kW = kW;
  g.outW = (inW * upx + padx0 + padx1 - kW + downx) / downx;
  g.outH = (inH * upy + pady0 + pady1 - kH + downy) / downy;
  TBG_CHECK_ARG(g.outW >= 1 && g.outH >= 1, "output must be at least 1x1");
  TBG_CHECK_ARG(major <= 65535, "tbg_upfirdn2d: major dimension too large (%d)", major);
  // input patch of an 8 x 32 output tile: samples between the first and the last upsampled coordinate it reads
  g.tile_rows = ((kUfTY - 1) * downy + kH - 1) / upy + 2;
  g.tile_cols = ((kUfTX - 1) * downx + kW - 1) / upx + 2;
  g.mchunk = minor < 8 ? minor : 8;
  size_t smem = (static_cast<size_t>(kH) * kW + static_cast<size_t>(g.tile_rows) * g.tile_cols * g.mchunk) * sizeof(float);
  if (smem > 48 * 1024) {
    g.mchunk = 1;
    smem = (static_cast<size_t>(kH) * kW + static_cast<size_t>(g.tile_rows) * g.tile_cols) * sizeof(float);
  }
  TBG_CHECK_ARG(smem <= 48 * 1024, "tbg_upfirdn2d: filter / decimation too large for the staged tile (%zu bytes)", smem);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  dim3 grid((g.outW + kUfTX - 1) / kUfTX, (g.outH + kUfTY - 1) / kUfTY, major), block(kUfTX, kUfTY);
  TBG_CHECK_ARG(grid.y <= 65535, "tbg_upfirdn2d: output too tall");
  if (dtype_bf16)
    upfirdn2d_tiled_kernel<__nv_bfloat16><<<grid, block, smem, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), k,
                                                                        reinterpret_cast<__nv_bfloat16*>(y), g);
  else
    upfirdn2d_tiled_kernel<float><<<grid, block, smem, stream>>>(reinterpret_cast<const float*>(x), k,
                                                                reinterpret_cast<float*>(y), g);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_torgb_skip_fwd(const void* x, const float* ws, const float* bias, const float* y_prev, const int* words,
                                  float* out, int B, int H, int W, int C, int mcn, int nchw, void* stream_v) {
  TBG_CHECK_ARG(x && ws && out, "tbg_torgb_skip_fwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1, "tbg_torgb_skip_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  TBG_CHECK_ARG(C == 64 || C == 128 || C == 192 || C == 256 || C == 512,
                "tbg_torgb_skip_fwd: C=%d not supported (64, 128, 192, 256 or 512 channels)", C);
  TBG_CHECK_ARG(!y_prev || (H % 2 == 0 && W % 2 == 0), "tbg_torgb_skip_fwd: the skip input needs even H and W");
  TBG_CHECK_ARG(!words || mcn >= 1, "tbg_torgb_skip_fwd: mask needs max_char_number");
  TBG_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "tbg_torgb_skip_fwd: x must be 16-byte aligned");
  TBG_CHECK_ARG(B <= 65535, "tbg_torgb_skip_fwd: batch too large");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int hw = H * W;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const uint4* xv = reinterpret_cast<const uint4*>(x);
  // Runs of pixels per sample sized from the kernel's real occupancy: whole waves of resident CTAs (a 5 % third wave
  // cost a third of the launch), each CTA long enough (>= 64 pixels) to amortise its 24 weight loads per lane.
#define TBG_TORGB(LPP, CV)                                                                                              \
  do {                                                                                                                  \
    static int occ = 0;                                                                                                 \
    if (!occ) {                                                                                                         \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, torgb_skip_fwd_kernel<LPP, CV>, 256, 0);                      \
      if (occ < 1) occ = 1;                                                                                             \
    }                                                                                                                   \
    const int pix_per_cta = torgb_run(B, hw, occ * sms, (256 / LPP) * ((CV) == 1 ? 8 : ((CV) == 2 ? 4 : 2)));            \
    dim3 grid((hw + pix_per_cta - 1) / pix_per_cta, B);                                                                 \
    torgb_skip_fwd_kernel<LPP, CV><<<grid, 256, 0, stream>>>(xv, ws, bias, y_prev, words, out, H, W, mcn, nchw,         \
                                                             pix_per_cta);                                              \
  } while (0)
  switch (C) {
    case 64: TBG_TORGB(8, 1); break;
    case 128: TBG_TORGB(16, 1); break;
    case 192: TBG_TORGB(8, 3); break;
    case 256: TBG_TORGB(32, 1); break;
    default: TBG_TORGB(32, 2); break;
  }
#undef TBG_TORGB
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_image_grad_nhwc(const float* g, const int* words, float* out, int B, int H, int W, int mcn, void* stream_v) {
  TBG_CHECK_ARG(g && out && B >= 1 && H >= 1 && W >= 1, "tbg_image_grad_nhwc: bad arguments");
  TBG_CHECK_ARG(!words || mcn >= 1, "tbg_image_grad_nhwc: mask needs max_char_number");
  const long long total = static_cast<long long>(B) * H * W;
  const int blocks = static_cast<int>((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  image_grad_nhwc_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(g, words, out, B, H, W, mcn);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
