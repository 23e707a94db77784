// Fused HBM-bound passes around the tensor-core convolutions of the plain (first-order) training
// step.  Each replaces a chain of element-wise / reduction ops of the reference graph:
//
//   modulate          xs = x * s[b,c]                              (modulated_conv2d.py:96)
//   modulate_bwd      gx = gxs * s ; gs[b,c] = sum_hw gxs * x       (its gradient)
//   bias_act_bwd      gradient of  out = act(y0*d + noise*ns + bias)*gain (+ residual)
//                     (modulated_conv2d.py:121, noise.py:21, bias_act.py:25-34, discriminator.py:82):
//                     gy0 = g_pre*d and the three per-(b,c) sums from which the gradients of
//                     d, bias and the noise strength follow
//   torgb_fwd / bwd   y[p,j] = sum_c x[p,c]*ws[b,c,j] (+ bias)      (to_rgb.py:28-33, N = 3 — HBM-bound)
//
// Layout: NHWC bf16 activations, 8 channels (16 bytes) per thread, fp32 math.  Reductions over
// pixels are done per CTA in registers and combined with fp32 atomics into [B, C] buffers.
#include <map>
#include <tuple>

#include "common.cuh"
#include "host_util.h"

namespace tbg {

static int sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

struct bf16x8 {
  uint4 v;
};
// bf16 -> fp32 is a 16-bit shift: one instruction per element (shift for the low half, mask for the high half) instead
// of the byte-permute + shift pair the bf162 intrinsic compiles to; these kernels are instruction-issue bound
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16);
  f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16);
  f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ---------------------------------------------------------------------------------------------
// xs = x * s[b, c]
// ---------------------------------------------------------------------------------------------
__global__ void modulate_kernel(const uint4* __restrict__ x, const float* __restrict__ s, uint4* __restrict__ xs,
                                long long n_vec, int hw, int c8) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const int cv = static_cast<int>(i % c8);
    const long long pix = i / c8;
    const int b = static_cast<int>(pix / hw);
    float f[8];
    unpack8(__ldg(x + i), f);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + (static_cast<long long>(b) * c8 + cv) * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + (static_cast<long long>(b) * c8 + cv) * 8 + 4));
    f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
    f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
    xs[i] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// gx = gxs * s ; gs[b,c] += sum_{pixels of the CTA's chunk} gxs * x
// grid = (chunks, B); block = c8 * rows threads: thread (r, cv) walks pixels r, r+rows, ...
// ---------------------------------------------------------------------------------------------
// out = act(t + noise[b,p]*ns + bias[c]) * gain — the element-wise tail of a layer (noise.py:21, bias_act.py:25-34) as a
// stand-alone pass; the plain training step applies it in the convolution epilogue, the twice-differentiable path of the
// regularisers (path length / R1) needs it as a separate differentiable node.
__global__ void bias_act_fwd_kernel(const uint4* __restrict__ t, const float* __restrict__ noise,
                                    const float* __restrict__ ns, const float* __restrict__ bias, uint4* __restrict__ out,
                                    long long n_vec, int hw, int c8, int act, float gain) {
  const float nsv = (noise != nullptr) ? __ldg(ns) : 0.f;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const int cv = static_cast<int>(i % c8);
    const long long pix = i / c8;
    float f[8];
    unpack8(__ldg(t + i), f);
    const float nz = (noise != nullptr) ? __ldg(noise + pix) * nsv : 0.f;
    float bv[8];
    if (bias != nullptr) {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cv * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cv * 8 + 4));
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) bv[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = f[k] + nz + bv[k];
      if (act == 1) v = v > 0.f ? v : 0.2f * v;
      else if (act == 2) v = fmaxf(v, 0.f);
      f[k] = v * gain;
    }
    out[i] = pack8(f);
  }
}

__global__ void modulate_bwd_kernel(const uint4* __restrict__ gxs, const uint4* __restrict__ x,
                                    const float* __restrict__ s, uint4* __restrict__ gx, float* __restrict__ gs,
                                    int hw, int c8, int pix_per_cta) {
  extern __shared__ float red[];  // [rows][c8*8]
  const int b = blockIdx.y;
  const int rows = blockDim.x / c8;
  const int cv = threadIdx.x % c8;
  const int r = threadIdx.x / c8;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  float sv[8], acc[8];
  if (gx != nullptr) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(s + (static_cast<long long>(b) * c8 + cv) * 8));
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(s + (static_cast<long long>(b) * c8 + cv) * 8 + 4));
    sv[0] = s0.x; sv[1] = s0.y; sv[2] = s0.z; sv[3] = s0.w; sv[4] = s1.x; sv[5] = s1.y; sv[6] = s1.z; sv[7] = s1.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (r < rows) {
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += rows) {
      const long long idx = (static_cast<long long>(b) * hw + p) * c8 + cv;
      float g[8], xv[8];
      unpack8(__ldg(gxs + idx), g);
      unpack8(__ldg(x + idx), xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(g[i], xv[i], acc[i]);
      if (gx != nullptr) {              // gx == nullptr: reduction only (tbg_rowdot)
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] *= sv[i];
        gx[idx] = pack8(g);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[(r * c8 + cv) * 8 + i] = acc[i];
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += red[(rr * c8 + cv) * 8 + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(gs + (static_cast<long long>(b) * c8 + cv) * 8 + i, acc[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// Backward of  out = act(y0*d + nz*ns + bias)*gain (+ res):
//   slope = act ? (out - res > 0 ? 1 : 0.2) : 1 ;  g_pre = g_out*gain*slope ;  gy0 = g_pre*d
//   S1[b,c] += g_pre ; Spre[b,c] += g_pre*pre ; Snz[b,c] += g_pre*nz      (pre = (out-res)/(gain*slope))
// ---------------------------------------------------------------------------------------------
// MODE 0: gy0 only; 1: + S1 (bias gradient); 2: + S1, Spre, Snz (demodulation / noise-strength gradients).
// Written multiply-only (the slope and its inverse are selected, not divided): the kernel is bandwidth-bound
// (3 activation passes) only when the per-element arithmetic stays below ~10 instructions.
template <int MODE>
__global__ void __launch_bounds__(256, 3)
bias_act_bwd_kernel(const uint4* __restrict__ g_out, const uint4* __restrict__ out,
                    const uint4* __restrict__ res, const float* __restrict__ noise,
                    const float* __restrict__ d, uint4* __restrict__ gy0, float* __restrict__ S1,
                    float* __restrict__ Spre, float* __restrict__ Snz, int hw, int c8,
                    int pix_per_cta, int act, float gain, int s1_over_batch) {
  extern __shared__ float red[];  // [rows][c8*8][MODE == 2 ? 3 : 1]
  const int b = blockIdx.y;
  const int rows = blockDim.x / c8;
  const int cv = threadIdx.x % c8;
  const int r = threadIdx.x / c8;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  float dv[8], a1[8], a2[8], a3[8];
  if (d != nullptr) {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(d + (static_cast<long long>(b) * c8 + cv) * 8));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(d + (static_cast<long long>(b) * c8 + cv) * 8 + 4));
    dv[0] = d0.x; dv[1] = d0.y; dv[2] = d0.z; dv[3] = d0.w; dv[4] = d1.x; dv[5] = d1.y; dv[6] = d1.z; dv[7] = d1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) dv[i] = 1.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) a1[i] = a2[i] = a3[i] = 0.f;
  // slope of the negative side and the factor that recovers the pre-activation from the output there
  const float neg_slope = act == 0 ? 1.f : (act == 2 ? 0.f : 0.2f);
  const float g_pos = gain, g_neg = gain * neg_slope;                     // g_pre = g_out * (out > 0 ? g_pos : g_neg)
  const float inv_gain = 1.f / gain;
  const float r_pos = inv_gain;                                           // pre = out / gain on the positive side
  if (r < rows) {
#pragma unroll 4
    for (int p = p0 + r; p < p1; p += rows) {
      const long long pix = static_cast<long long>(b) * hw + p;
      const long long idx = pix * c8 + cv;
      float g[8], o[8];
      unpack8(__ldg(g_out + idx), g);
      unpack8(__ldg(out + idx), o);
      if (res != nullptr) {
        float rv[8];
        unpack8(__ldg(res + idx), rv);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] -= rv[i];
      }
      float nz = 0.f;
      if (MODE == 2 && noise != nullptr) nz = __ldg(noise + pix);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const bool pos = o[i] > 0.f;
        const float gp = g[i] * (pos ? g_pos : g_neg);
        if (MODE >= 1) a1[i] += gp;
        if (MODE == 2) {
          // g_pre * pre: the slope and its inverse cancel (linear / leaky-ReLU), so no second select is needed
          a2[i] = fmaf(act == 2 ? gp * r_pos : g[i], o[i], a2[i]);
          a3[i] = fmaf(gp, nz, a3[i]);
        }
        g[i] = gp * dv[i];
      }
      gy0[idx] = pack8(g);
    }
  }
  if (MODE == 0) return;
  constexpr int NS = MODE == 2 ? 3 : 1;
  const int cw = c8 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    red[(r * cw + cv * 8 + i) * NS + 0] = a1[i];
    if (MODE == 2) {
      red[(r * cw + cv * 8 + i) * NS + 1] = a2[i];
      red[(r * cw + cv * 8 + i) * NS + 2] = a3[i];
    }
  }
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a1[i] += red[(rr * cw + cv * 8 + i) * NS + 0];
        if (MODE == 2) {
          a2[i] += red[(rr * cw + cv * 8 + i) * NS + 1];
          a3[i] += red[(rr * cw + cv * 8 + i) * NS + 2];
        }
      }
    const long long o0 = (static_cast<long long>(s1_over_batch ? 0 : b) * c8 + cv) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(S1 + o0 + i, a1[i]);
      if (MODE == 2) {
        if (Spre != nullptr) atomicAdd(Spre + o0 + i, a2[i]);
        if (noise != nullptr) atomicAdd(Snz + o0 + i, a3[i]);
      }
    }
  }
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  // 16-byte global -> shared copy that bypasses registers; src-size 0 writes zeros (out-of-bounds taps)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 16 : 0)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// bias_act_bwd of a modulated layer whose output also feeds a ToRGB (synthesis_block.py:143-152): the gradient that
// reaches `out` is g_out (from the next block, absent on the last one) + g_rgb (x) ws — a rank-3 term that is formed here
// from the [B,HW,3] image gradient instead of being materialised, added and re-read — and the ToRGB weight gradient
// gws[b,c,j] = sum_p out[b,p,c] * g_rgb[b,p,j] needs the same pass over `out`:
//   g = (g_out + sum_j g_rgb[p,j] ws[b,c,j]);  then exactly MODE 2 of bias_act_bwd_kernel (S1, Spre, Snz, gy0 = g_pre*d).
// ---------------------------------------------------------------------------------------------
template <bool HAS_G>
__global__ void __launch_bounds__(256, 2)
bias_act_rgb_bwd_kernel(const uint4* __restrict__ g_out, const uint4* __restrict__ out, const float* __restrict__ noise,
                        const float* __restrict__ d, const float* __restrict__ g_rgb, const float* __restrict__ ws,
                        uint4* __restrict__ gy0, float* __restrict__ S1, float* __restrict__ Spre,
                        float* __restrict__ Snz, float* __restrict__ gws, int hw, int c8, int pix_per_cta, int act,
                        float gain) {
  extern __shared__ float red[];  // [rows][c8*8][6]
  const int b = blockIdx.y;
  const int rows = blockDim.x / c8;
  const int cv = threadIdx.x % c8;
  const int r = threadIdx.x / c8;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  float dv[8], w[8][3], a1[8], a2[8], a3[8], aw[8][3];
  {
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(d + (static_cast<long long>(b) * c8 + cv) * 8));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(d + (static_cast<long long>(b) * c8 + cv) * 8 + 4));
    dv[0] = d0.x; dv[1] = d0.y; dv[2] = d0.z; dv[3] = d0.w; dv[4] = d1.x; dv[5] = d1.y; dv[6] = d1.z; dv[7] = d1.w;
    const float* wp = ws + (static_cast<long long>(b) * c8 + cv) * 24;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a1[i] = a2[i] = a3[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        w[i][j] = __ldg(wp + i * 3 + j);
        aw[i][j] = 0.f;
      }
    }
  }
  const float neg_slope = act == 0 ? 1.f : (act == 2 ? 0.f : 0.2f);
  const float g_pos = gain, g_neg = gain * neg_slope;
  const float r_pos = 1.f / gain;
  // The two 16-byte streams (g_out, out) go through a thread-private cp.async ring, kRgbStages - 1 pixels ahead: the
  // 80 accumulator / weight registers leave room for two CTAs of 8 warps per SM only, and with the loads held in
  // registers (one or two pixels per thread) that is ~32 KB in flight per SM, 0.53 of the HBM rate.  The ring lives in
  // the shared memory of the final reduction (same size: 6 floats x 8 channels = 2 x 16 B x 6 stages per thread).
  constexpr int kRgbStages = 6;
  uint4* ring = reinterpret_cast<uint4*>(red);
  const int nthr = blockDim.x, tid = threadIdx.x;
  const long long pix0 = static_cast<long long>(b) * hw;
  auto issue = [&](int p, int slot) {
    const bool ok = p < p1;
    const long long idx = (pix0 + (ok ? p : p0)) * c8 + cv;
    if (HAS_G) cp_async_16(ring + (slot * 2 + 0) * nthr + tid, g_out + idx, ok);
    cp_async_16(ring + (slot * 2 + 1) * nthr + tid, out + idx, ok);
    cp_async_commit();
  };
#pragma unroll
  for (int st = 0; st < kRgbStages - 1; ++st) issue(p0 + r + st * rows, st);
  {
    int slot = 0, slot_new = kRgbStages - 1;
    // per-pixel scalars (image gradient, noise) one pixel ahead in registers
    float n0 = 0.f, n1 = 0.f, n2 = 0.f, nnz = 0.f;
    if (p0 + r < p1) {
      const long long pix = pix0 + p0 + r;
      n0 = __ldg(g_rgb + pix * 3 + 0); n1 = __ldg(g_rgb + pix * 3 + 1); n2 = __ldg(g_rgb + pix * 3 + 2);
      if (noise != nullptr) nnz = __ldg(noise + pix);
    }
    for (int p = p0 + r; p < p1; p += rows) {
      issue(p + (kRgbStages - 1) * rows, slot_new);
      const float r0 = n0, r1 = n1, r2 = n2, nz = nnz;
      if (p + rows < p1) {
        const long long pixn = pix0 + p + rows;
        n0 = __ldg(g_rgb + pixn * 3 + 0); n1 = __ldg(g_rgb + pixn * 3 + 1); n2 = __ldg(g_rgb + pixn * 3 + 2);
        if (noise != nullptr) nnz = __ldg(noise + pixn);
      }
      cp_async_wait<kRgbStages - 1>();
      float g[8], o[8];
      if (HAS_G) {
        unpack8(ring[(slot * 2 + 0) * nthr + tid], g);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
      }
      unpack8(ring[(slot * 2 + 1) * nthr + tid], o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        aw[i][0] = fmaf(o[i], r0, aw[i][0]);
        aw[i][1] = fmaf(o[i], r1, aw[i][1]);
        aw[i][2] = fmaf(o[i], r2, aw[i][2]);
        const float gt = g[i] + (r0 * w[i][0] + r1 * w[i][1] + r2 * w[i][2]);
        const bool pos = o[i] > 0.f;
        const float gp = gt * (pos ? g_pos : g_neg);
        a1[i] += gp;
        a2[i] = fmaf(act == 2 ? gp * r_pos : gt, o[i], a2[i]);
        a3[i] = fmaf(gp, nz, a3[i]);
        g[i] = gp * dv[i];
      }
      gy0[(pix0 + p) * c8 + cv] = pack8(g);
      slot = slot + 1 == kRgbStages ? 0 : slot + 1;
      slot_new = slot_new + 1 == kRgbStages ? 0 : slot_new + 1;
    }
  }
  cp_async_wait<0>();
  __syncthreads();                          // the reduction buffer below aliases every thread's ring slots
  const int cw = c8 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float* rp = red + (r * cw + cv * 8 + i) * 6;
    rp[0] = a1[i]; rp[1] = a2[i]; rp[2] = a3[i];
    rp[3] = aw[i][0]; rp[4] = aw[i][1]; rp[5] = aw[i][2];
  }
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float* rp = red + (rr * cw + cv * 8 + i) * 6;
        a1[i] += rp[0]; a2[i] += rp[1]; a3[i] += rp[2];
        aw[i][0] += rp[3]; aw[i][1] += rp[4]; aw[i][2] += rp[5];
      }
    const long long o0 = (static_cast<long long>(b) * c8 + cv) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(S1 + o0 + i, a1[i]);
      atomicAdd(Spre + o0 + i, a2[i]);
      if (noise != nullptr) atomicAdd(Snz + o0 + i, a3[i]);
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(gws + (o0 + i) * 3 + j, aw[i][j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// ToRGB: y[p, j] = sum_c x[p,c] * ws[b,c,j] (+ bias[j]); one warp per pixel, lanes over channels.
// ---------------------------------------------------------------------------------------------
__global__ void torgb_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ ws, const float* __restrict__ bias,
                                 float* __restrict__ y, int B, int hw, int c8) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const long long n_pix = static_cast<long long>(B) * hw;
  for (long long pix = static_cast<long long>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5); pix < n_pix;
       pix += static_cast<long long>(gridDim.x) * warps_per_block) {
    const int b = static_cast<int>(pix / hw);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int cv = lane; cv < c8; cv += 32) {
      float f[8];
      unpack8(__ldg(x + pix * c8 + cv), f);
      const float* w = ws + (static_cast<long long>(b) * c8 + cv) * 24;  // [c][3]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a0 = fmaf(f[i], __ldg(w + i * 3 + 0), a0);
        a1 = fmaf(f[i], __ldg(w + i * 3 + 1), a1);
        a2 = fmaf(f[i], __ldg(w + i * 3 + 2), a2);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
      float* yo = y + pix * 3;
      yo[0] = a0 + (bias ? __ldg(bias + 0) : 0.f);
      yo[1] = a1 + (bias ? __ldg(bias + 1) : 0.f);
      yo[2] = a2 + (bias ? __ldg(bias + 2) : 0.f);
    }
  }
}

// gx[p,c] = sum_j gy[p,j]*ws[b,c,j] ; gws[b,c,j] += sum_p x[p,c]*gy[p,j]
__global__ void torgb_bwd_kernel(const uint4* __restrict__ x, const float* __restrict__ ws, const float* __restrict__ gy,
                                 uint4* __restrict__ gx, float* __restrict__ gws, int hw, int c8, int pix_per_cta) {
  extern __shared__ float red[];  // [rows][c8*8][3]
  const int b = blockIdx.y;
  const int rows = blockDim.x / c8;
  const int cv = threadIdx.x % c8;
  const int r = threadIdx.x / c8;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  float w[8][3], acc[8][3];
  {
    const float* wp = ws + (static_cast<long long>(b) * c8 + cv) * 24;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        w[i][j] = __ldg(wp + i * 3 + j);
        acc[i][j] = 0.f;
      }
  }
  if (r < rows) {
#pragma unroll 2
    for (int p = p0 + r; p < p1; p += rows) {
      const long long pix = static_cast<long long>(b) * hw + p;
      const float g0 = __ldg(gy + pix * 3 + 0), g1 = __ldg(gy + pix * 3 + 1), g2 = __ldg(gy + pix * 3 + 2);
      float xv[8], o[8];
      unpack8(__ldg(x + pix * c8 + cv), xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = g0 * w[i][0] + g1 * w[i][1] + g2 * w[i][2];
        acc[i][0] = fmaf(xv[i], g0, acc[i][0]);
        acc[i][1] = fmaf(xv[i], g1, acc[i][1]);
        acc[i][2] = fmaf(xv[i], g2, acc[i][2]);
      }
      gx[pix * c8 + cv] = pack8(o);
    }
  }
  const int cw = c8 * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) red[(r * cw + cv * 8 + i) * 3 + j] = acc[i][j];
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[i][j] += red[(rr * cw + cv * 8 + i) * 3 + j];
    float* o = gws + (static_cast<long long>(b) * c8 + cv) * 24;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) atomicAdd(o + i * 3 + j, acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// 4x4 separable FIR [1,3,3,1] x [1,3,3,1] on NHWC bf16 (the resample kernel of synthesis_block.py:36 /
// discriminator.py:44 as applied by upfirdn_2d, upfirdn_2d_v2.py:116-163):
//   out[b,y,x,c] = scale * sum_{m,n<4} k[m] k[n] in[b, y+m+offy, x+n+offx, c]     (in = 0 out of bounds)
// optionally followed by the layer epilogue  v = act(v*d[b,c] + noise[b,y,x]*ns + bias[c]) * gain.
// The kernel is symmetric, so the adjoint is the same call with off' = -3 - off and in/out swapped.
// ---------------------------------------------------------------------------------------------
static constexpr int kFirCols = 32, kFirStrip = 32;   // CTA: 32 output columns x 8 channel vectors, a strip of 32 rows
static constexpr int kFirRing = 8, kFirAhead = 6;     // fir4_kernel: input rows staged in shared memory / copies in flight

// 4 x 4 FIR [1,3,3,1] x [1,3,3,1] * scale on NHWC bf16 with the optional layer epilogue.
// Thread (x, 8-channel vector) walks down a strip of output rows keeping the last four horizontally filtered rows in
// registers.  The round-2 version loaded the four horizontal taps of a row straight from global memory: three of the
// four are L1 hits, so a warp had only 512 UNIQUE bytes in flight per row and the kernel sat at 0.48 of HBM, bound by
// memory latency (profiles/r02ai_perf_pointwise.log).  Now every input row of the CTA's tile (35 pixels x 64 channels) is
// copied ONCE with cp.async into a shared-memory ring, six rows ahead of the row being filtered (zero-filled outside the
// tensor), and the taps are conflict-free 16-byte shared-memory reads.
// EPI = false: bare filter (the FIR adjoint of the up layers' backward pass, the blur in front of the strided
// discriminator convolutions) with a smaller register footprint.
template <bool EPI>
__global__ void __launch_bounds__(256)
fir4_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int IH, int IW, int OH, int OW, int c8, int offy,
            int offx, float scale, const float* __restrict__ d, const float* __restrict__ noise,
            const float* __restrict__ ns, const float* __restrict__ bias, int act, float gain, int cgroups) {
  __shared__ uint4 ring[kFirRing][kFirCols + 3][8];
  const int x0 = blockIdx.x * kFirCols, y0 = blockIdx.y * kFirStrip;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int v = threadIdx.x & 7, xl = threadIdx.x >> 3;
  const int x = x0 + xl, cv = cg * 8 + v;
  const bool mine = x < OW && cv < c8;                     // this thread owns an output column (all threads copy / sync)
  const uint4* src = in + static_cast<long long>(b) * IH * IW * c8;
  const int rows_out = min(kFirStrip, OH - y0);
  const int rows_in = rows_out + 3;
  // input row r of the tile (tensor row y0 + offy + r) -> ring slot r % kFirRing; one commit group per row, always.
  // Every thread copies pixel column xl (and threads 0..23 also one of the three halo columns 32..34).  The kernel is
  // issue-bound (ncu: 0.89 instructions per scheduler cycle), so everything row-invariant is hoisted and the per-row
  // state (source pointers, tensor row, ring slot) is carried incrementally: rows are fetched strictly in order.
  const int ix_a = x0 + offx + xl, ix_b = ix_a + kFirCols;
  const bool ok_a = cv < c8 && ix_a >= 0 && ix_a < IW;
  const bool has_b = xl < 3;
  const bool ok_b = has_b && cv < c8 && ix_b >= 0 && ix_b < IW;
  const long long row_stride = static_cast<long long>(IW) * c8;
  int f_iy = y0 + offy;                                    // tensor row of the next fetch
  int f_left = rows_in;                                    // rows still to fetch
  const uint4* f_pa = (ok_a ? src + static_cast<long long>(ix_a) * c8 + cv : src) + static_cast<long long>(f_iy) * row_stride;
  const uint4* f_pb = (ok_b ? src + static_cast<long long>(ix_b) * c8 + cv : src) + static_cast<long long>(f_iy) * row_stride;
  constexpr uint32_t kSlotBytes = (kFirCols + 3) * 8 * 16;
  const uint32_t sa0 = smem_u32(&ring[0][xl][v]), sb0 = smem_u32(&ring[0][kFirCols + (xl % 3)][v]);
  uint32_t f_so = 0;                                       // byte offset of the ring slot of the next fetch
  auto fetch = [&]() {
    if (f_left > 0) {
      const bool row_ok = f_iy >= 0 && f_iy < IH;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa0 + f_so), "l"(row_ok ? f_pa : src),
                   "r"(ok_a && row_ok ? 16 : 0)
                   : "memory");
      if (has_b)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sb0 + f_so), "l"(row_ok ? f_pb : src),
                     "r"(ok_b && row_ok ? 16 : 0)
                     : "memory");
    }
    cp_async_commit();
    --f_left;
    ++f_iy;
    f_pa += row_stride;
    f_pb += row_stride;
    f_so = (f_so == (kFirRing - 1) * kSlotBytes) ? 0u : f_so + kSlotBytes;
  };
  // horizontally filtered input row, from the ring slot at byte offset h_so (rows are consumed strictly in order):
  // (t0 + t3) + 3 (t1 + t2)
  uint32_t h_so = 0;
  const uint8_t* const ring_mine = reinterpret_cast<const uint8_t*>(&ring[0][xl][v]);
  auto hrow = [&](float (&h)[8]) {
    const uint4* row = reinterpret_cast<const uint4*>(ring_mine + h_so);
    float f0[8], f1[8], f2[8], f3[8];
    unpack8(row[0], f0);
    unpack8(row[8], f1);
    unpack8(row[16], f2);
    unpack8(row[24], f3);
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = fmaf(3.f, f1[i] + f2[i], f0[i] + f3[i]);
    h_so = (h_so == (kFirRing - 1) * kSlotBytes) ? 0u : h_so + kSlotBytes;
  };
  float dv[8], bv[8];
  float nsv = 0.f;
  if (EPI) {
    nsv = (noise != nullptr) ? __ldg(ns) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dv[i] = (d != nullptr && cv < c8) ? __ldg(d + (static_cast<long long>(b) * c8 + cv) * 8 + i) * scale : scale;
      bv[i] = (bias != nullptr && cv < c8) ? __ldg(bias + cv * 8 + i) : 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < kFirAhead; ++r) fetch();
  float win[4][8];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    cp_async_wait<kFirAhead - 1>();
    __syncthreads();
    hrow(win[r]);
    fetch();
  }
  uint4* optr = out + ((static_cast<long long>(b) * OH + y0) * OW + x) * c8 + cv;     // output row y0 of this thread
  const long long out_stride = static_cast<long long>(OW) * c8;
  const float* nptr = (EPI && noise != nullptr) ? noise + (static_cast<long long>(b) * OH + y0) * OW + x : nullptr;
  asm volatile("" : "+l"(optr));           // opaque to the optimiser: keep it in a register pair
  for (int yb = 0; yb < rows_out; yb += 4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {          // q is a compile-time constant: the ring indices below stay in registers
      const int yy = yb + q;
      if (yy >= rows_out) break;           // uniform over the CTA
      cp_async_wait<kFirAhead - 1>();
      __syncthreads();
      hrow(win[(q + 3) & 3]);
      fetch();
      uint4* const o_row = optr;             // carried incrementally: the compiler otherwise re-derives the 64-bit
      const float* const n_row = nptr;       // address from the block indices every row (~25 instructions)
      optr += out_stride;
      if (nptr != nullptr) nptr += OW;
      if (!mine) continue;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        acc[i] = win[q & 3][i] + win[(q + 3) & 3][i] + 3.f * (win[(q + 1) & 3][i] + win[(q + 2) & 3][i]);
      if (EPI) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= dv[i];
        if (n_row != nullptr) {
          const float nz = __ldg(n_row) * nsv;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += nz;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float o = acc[i] + bv[i];
          if (act == 1) o = o > 0.f ? o : 0.2f * o;
          acc[i] = o * gain;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= scale;
      }
      *o_row = pack8(acc);
    }
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// The same 4x4 FIR evaluated only at the pixels a stride-(SY, 2) 1x1 convolution reads (the skip branch of a residual
// block, discriminator.py:126-131 -> conv_downsample_2d with a 1x1 kernel, upfirdn_2d_v2.py:106-113):
//   out[b,p,q,c] = scale * sum_{m,n<4} k[m] k[n] in[b, SY*p+m+offy, 2*q+n+offx, c]        (in = 0 out of bounds)
// Same thread layout as fir4_kernel; with SY = 2 two new horizontally filtered rows enter the window per output row.
// ---------------------------------------------------------------------------------------------
template <int SY>
__global__ void __launch_bounds__(256)
fir4_down_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int IH, int IW, int OH, int OW, int c8, int offy,
                 int offx, float scale, int cgroups) {
  const int x0 = blockIdx.x * kFirCols, y0 = blockIdx.y * kFirStrip;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int v = threadIdx.x & 7, xl = threadIdx.x >> 3;
  const int x = x0 + xl, cv = cg * 8 + v;
  if (x >= OW || cv >= c8) return;
  const uint4* src = in + static_cast<long long>(b) * IH * IW * c8 + cv;
  const int ix0 = 2 * x + offx;
  auto hrow = [&](int iy, float (&h)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) h[i] = 0.f;
    if (iy < 0 || iy >= IH) return;
    const uint4* row = src + static_cast<long long>(iy) * IW * c8;
    uint4 t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ix = ix0 + k;
      t[k] = (ix >= 0 && ix < IW) ? __ldg(row + static_cast<long long>(ix) * c8) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float f[8];
      unpack8(t[k], f);
      const float wk = (k == 0 || k == 3) ? 1.f : 3.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) h[i] = fmaf(f[i], wk, h[i]);
    }
  };
  const int y_end = min(y0 + kFirStrip, OH);
  uint4* dst = out + (static_cast<long long>(b) * OH * OW + x) * c8 + cv;
  float win[4][8];
  if (SY == 1) {
    hrow(y0 + offy + 0, win[0]);
    hrow(y0 + offy + 1, win[1]);
    hrow(y0 + offy + 2, win[2]);
    for (int yb = y0; yb < y_end; yb += 4) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int y = yb + q;
        if (y >= y_end) return;
        hrow(y + offy + 3, win[(q + 3) & 3]);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          acc[i] = (win[q & 3][i] + win[(q + 3) & 3][i] + 3.f * (win[(q + 1) & 3][i] + win[(q + 2) & 3][i])) * scale;
        dst[static_cast<long long>(y) * OW * c8] = pack8(acc);
      }
    }
  } else {
    hrow(2 * y0 + offy + 0, win[0]);
    hrow(2 * y0 + offy + 1, win[1]);
    for (int yb = y0; yb < y_end; yb += 2) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {          // q compile-time: the window halves swap roles without register moves
        const int y = yb + q;
        if (y >= y_end) return;
        const int lo = 2 * q, hi = 2 * (q ^ 1);
        hrow(2 * y + offy + 2, win[hi]);
        hrow(2 * y + offy + 3, win[hi + 1]);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          acc[i] = (win[lo][i] + win[hi + 1][i] + 3.f * (win[lo + 1][i] + win[hi][i])) * scale;
        dst[static_cast<long long>(y) * OW * c8] = pack8(acc);
      }
    }
  }
}

// Adjoint of fir4_down (the input gradient of the skip branch), added to the gradient arriving through the main branch:
//   gx[b,y,x,c] = add[b,y,x,c] + scale * sum k[m] k[n] g[b,p,q,c]   over (m,p): SY*p+m+offy = y and (n,q): 2*q+n+offx = x
// One thread per (x, 8-channel vector) walking down a strip of rows; the (4/SY) x 2 contributing elements of the
// (SY*2 times smaller) gradient tile are re-read through L1.
template <int SY>
__global__ void __launch_bounds__(256)
fir4_down_adjoint_kernel(const uint4* __restrict__ g, const uint4* __restrict__ add, uint4* __restrict__ out, int IH, int IW,
                         int OH, int OW, int c8, int offy, int offx, float scale, int cgroups) {
  const int x0 = blockIdx.x * kFirCols, y0 = blockIdx.y * kFirStrip;
  const int b = blockIdx.z / cgroups, cg = blockIdx.z % cgroups;
  const int v = threadIdx.x & 7, xl = threadIdx.x >> 3;
  const int x = x0 + xl, cv = cg * 8 + v;
  if (x >= IW || cv >= c8) return;
  const uint4* src = g + static_cast<long long>(b) * OH * OW * c8 + cv;
  // the two horizontal taps: n = n0, n0 + 2 with q = (x - offx - n) / 2
  const int n0 = (x - offx) & 1;
  const int q0 = (x - offx - n0) >> 1, q1 = q0 - 1;           // taps n0 and n0 + 2
  const float wq0 = n0 == 0 ? 1.f : 3.f, wq1 = n0 == 0 ? 3.f : 1.f;
  const bool ok0 = q0 >= 0 && q0 < OW, ok1 = q1 >= 0 && q1 < OW;
  const int y_end = min(y0 + kFirStrip, IH);
  for (int y = y0; y < y_end; ++y) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int t = y - offy - m;
      if (SY == 2 && (t & 1)) continue;
      const int p = SY == 2 ? (t >> 1) : t;
      if (t < 0 || p >= OH) continue;
      const float wm = (m == 0 || m == 3) ? 1.f : 3.f;
      const uint4* row = src + static_cast<long long>(p) * OW * c8;
      float f[8];
      if (ok0) {
        unpack8(__ldg(row + static_cast<long long>(q0) * c8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wm * wq0, acc[i]);
      }
      if (ok1) {
        unpack8(__ldg(row + static_cast<long long>(q1) * c8), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(f[i], wm * wq1, acc[i]);
      }
    }
    const long long o = ((static_cast<long long>(b) * IH + y) * IW + x) * c8 + cv;
    if (add != nullptr) {
      float f[8];
      unpack8(__ldg(add + o), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], scale, f[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] *= scale;
    }
    out[o] = pack8(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// FromRGB (from_rgb.py:26-29): 1x1 convolution 3 -> C on the NCHW fp32 image + bias + leaky-ReLU*gain,
// written as NHWC bf16.  K = 3 makes it bandwidth-bound: one thread per (pixel, 8 channels).
//   out[b,p,c] = lrelu(coef * sum_j img[b,j,p] * w[j,c] + bias[c]) * gain
// backward: gpre = g*gain*slope(out);  gimg[b,j,p] = coef * sum_c gpre*w[j,c];
//           gw[j,c] += coef * sum_{b,p} img*gpre;  gb[c] += sum_{b,p} gpre            (gw, gb zeroed by the caller)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fromrgb_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                   uint4* __restrict__ out, int hw, int c8, float coef, float gain, long long n_vec) {
  for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_vec;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(e % c8);
    const long long pix = e / c8;
    const long long b = pix / hw, p = pix - b * hw;
    const float* ip = img + b * 3 * hw + p;
    const float x0 = __ldg(ip) * coef, x1 = __ldg(ip + hw) * coef, x2 = __ldg(ip + 2 * static_cast<long long>(hw)) * coef;
    const int C = c8 * 8;
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cv * 8 + i;
      float v = fmaf(x0, __ldg(w + c), fmaf(x1, __ldg(w + C + c), fmaf(x2, __ldg(w + 2 * C + c), __ldg(bias + c))));
      v = v > 0.f ? v : 0.2f * v;
      f[i] = v * gain;
    }
    out[e] = pack8(f);
  }
}

// Same op when C / 8 divides the block (every ladder of the reference): a thread keeps ONE channel vector for all its
// pixels, so the 24 weights and 8 biases sit in registers — the generic kernel above re-reads them for every 16-byte
// store (35 loads per store, 0.22 of the HBM rate) — and per pixel only the three image values are loaded (shared by the
// c8 lanes of the pixel, consecutive pixels in consecutive lane groups).
__global__ void __launch_bounds__(256)
fromrgb_fwd_cv_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                      uint4* __restrict__ out, int hw, int c8, float coef, float gain, int pix_per_cta) {
  // grid (chunks, B): no per-pixel division
  const int cv = threadIdx.x % c8, r = threadIdx.x / c8, rows = blockDim.x / c8;
  const int C = c8 * 8;
  float wv[3][8], bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cv * 8 + i;
    wv[0][i] = __ldg(w + c) * coef; wv[1][i] = __ldg(w + C + c) * coef; wv[2][i] = __ldg(w + 2 * C + c) * coef;
    bv[i] = __ldg(bias + c);
  }
  const float g_pos = gain, g_neg = 0.2f * gain;
  const int p0 = blockIdx.x * pix_per_cta, p1 = min(p0 + pix_per_cta, hw);
  const float* ip = img + static_cast<long long>(blockIdx.y) * 3 * hw;
  uint4* op = out + static_cast<long long>(blockIdx.y) * hw * c8 + cv;
#pragma unroll 4
  for (int p = p0 + r; p < p1; p += rows) {
    const float x0 = __ldg(ip + p), x1 = __ldg(ip + hw + p), x2 = __ldg(ip + 2 * hw + p);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = fmaf(x0, wv[0][i], fmaf(x1, wv[1][i], fmaf(x2, wv[2][i], bv[i])));
      f[i] = v * (v > 0.f ? g_pos : g_neg);
    }
    op[static_cast<long long>(p) * c8] = pack8(f);
  }
}

// grid (chunks, B); block = c8 * rows threads (c8 <= 32, a power of two): thread (r, cv) walks pixels r, r+rows, ...
__global__ void fromrgb_bwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const uint4* __restrict__ g_out,
                                   const uint4* __restrict__ out, float* __restrict__ gimg, float* __restrict__ gw,
                                   float* __restrict__ gb, int hw, int c8, int pix_per_cta, float coef, float gain,
                                   int want_w) {
  extern __shared__ float red[];  // [rows][c8*8][4]
  const int b = blockIdx.y;
  const int rows = blockDim.x / c8;
  const int cv = threadIdx.x % c8;
  const int r = threadIdx.x / c8;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(p0 + pix_per_cta, hw);
  const int C = c8 * 8;
  float wv[3][8], acc[4][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) wv[j][i] = __ldg(w + j * C + cv * 8 + i) * coef;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][i] = 0.f;
  }
  // uniform trip count for every thread of the CTA: the per-pixel shuffle reduction below needs whole warps
#pragma unroll 2
  for (int pp = p0; pp < p1; pp += rows) {
    const int p = pp + r;
    const bool valid = (r < rows) && (p < p1);
    const long long pix = static_cast<long long>(b) * hw + (valid ? p : p0);
    float g[8], o[8];
    unpack8(__ldg(g_out + pix * c8 + cv), g);
    unpack8(__ldg(out + pix * c8 + cv), o);
    const float* ip = img + static_cast<long long>(b) * 3 * hw + (valid ? p : p0);
    const float x0 = __ldg(ip), x1 = __ldg(ip + hw), x2 = __ldg(ip + 2 * static_cast<long long>(hw));
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gp = valid ? g[i] * gain * (o[i] > 0.f ? 1.f : 0.2f) : 0.f;
      s0 = fmaf(gp, wv[0][i], s0);
      s1 = fmaf(gp, wv[1][i], s1);
      s2 = fmaf(gp, wv[2][i], s2);
      acc[0][i] = fmaf(gp, x0, acc[0][i]);
      acc[1][i] = fmaf(gp, x1, acc[1][i]);
      acc[2][i] = fmaf(gp, x2, acc[2][i]);
      acc[3][i] += gp;
    }
    // sum over the c8 threads of this pixel (consecutive lanes; c8 is a power of two <= 32)
    for (int sh = c8 >> 1; sh > 0; sh >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, sh);
      s1 += __shfl_xor_sync(0xffffffffu, s1, sh);
      s2 += __shfl_xor_sync(0xffffffffu, s2, sh);
    }
    if (valid && cv == 0 && gimg != nullptr) {
      float* gp = gimg + static_cast<long long>(b) * 3 * hw + p;
      gp[0] = s0;
      gp[hw] = s1;
      gp[2 * static_cast<long long>(hw)] = s2;
    }
  }
  if (!want_w) return;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[((r * c8 + cv) * 8 + i) * 4 + j] = acc[j][i];
  __syncthreads();
  if (r == 0) {
    for (int rr = 1; rr < rows; ++rr)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] += red[((rr * c8 + cv) * 8 + i) * 4 + j];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = cv * 8 + i;
      atomicAdd(gw + c, acc[0][i] * coef);
      atomicAdd(gw + C + c, acc[1][i] * coef);
      atomicAdd(gw + 2 * C + c, acc[2][i] * coef);
      atomicAdd(gb + c, acc[3][i]);
    }
  }
}

// launch geometry for the (chunks, B) reduction kernels
struct RedGeom {
  int threads, rows, pix_per_cta, chunks;
  size_t smem;
};
// Grid of the "stream over the pixels of a sample + reduce per (sample, channel)" kernels: (chunks, B) CTAs of rows x c8
// threads.  The chunk count is chosen against the number of CTAs the GPU holds at once for THIS kernel (occupancy query,
// cached): B * chunks just above one wave runs a second, nearly empty wave — with the former "~4 CTAs per SM" rule
// bias_act_bwd<2> at batch 64 launched 640 CTAs on 444 slots = 72 % wave efficiency (profiles/r02ai_perf_pointwise.log).
static RedGeom red_geom(int B, int hw, int c8, int per_elem_floats, const void* kernel) {
  RedGeom g;
  int rows = 256 / c8;
  if (rows < 1) rows = 1;
  if (rows > hw) rows = hw;
  g.rows = rows;
  g.threads = rows * c8;
  g.smem = static_cast<size_t>(rows) * c8 * 8 * per_elem_floats * sizeof(float);
  static std::map<std::tuple<const void*, int, size_t>, int> occ;
  const auto key = std::make_tuple(kernel, g.threads, g.smem);
  auto it = occ.find(key);
  if (it == occ.end()) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, g.threads, g.smem) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 2;
    }
    it = occ.emplace(key, n).first;
  }
  const long long slots = static_cast<long long>(it->second) * sms();
  const int max_chunks = (hw + rows - 1) / rows;
  int cmax = static_cast<int>((4 * slots + B - 1) / B);
  if (cmax > max_chunks) cmax = max_chunks;
  if (cmax < 1) cmax = 1;
  // cost model: waves x (pixel-row iterations of one CTA + its fixed cost: launch, shared-memory reduction, atomics ~ 16
  // iterations); more, smaller CTAs only pay while they fill the waves better
  int best = 1;
  double best_t = 1e300;
  for (int c = 1; c <= cmax; ++c) {
    const int ppc = (hw + c - 1) / c;
    const int cc = (hw + ppc - 1) / ppc;                   // chunks actually launched
    const long long total = static_cast<long long>(B) * cc;
    const long long waves = (total + slots - 1) / slots;
    const double t = static_cast<double>(waves) * (static_cast<double>((ppc + rows - 1) / rows) + 16.0);
    if (t < best_t - 1e-9) {
      best_t = t;
      best = c;
    }
  }
  g.pix_per_cta = (hw + best - 1) / best;
  g.chunks = (hw + g.pix_per_cta - 1) / g.pix_per_cta;
  return g;
}

}  // namespace tbg

using namespace tbg;

#define TBG_ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" int tbg_modulate(const void* x, const float* s, void* xs, int B, int HW, int C, void* stream_v) {
  TBG_CHECK_ARG(x && s && xs, "tbg_modulate: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && HW >= 1, "tbg_modulate: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(x) && TBG_ALIGNED16(s) && TBG_ALIGNED16(xs), "tbg_modulate: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n_vec = static_cast<long long>(B) * HW * (C / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = static_cast<long long>(sms()) * 16;
  if (blocks > cap) blocks = cap;
  modulate_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), s,
                                                                reinterpret_cast<uint4*>(xs), n_vec, HW, C / 8);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_modulate_bwd(const void* gxs, const void* x, const float* s, void* gx, float* gs, int B, int HW, int C,
                                void* stream_v) {
  TBG_CHECK_ARG(gxs && x && s && gx && gs, "tbg_modulate_bwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && B >= 1 && HW >= 1, "tbg_modulate_bwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(gxs) && TBG_ALIGNED16(x) && TBG_ALIGNED16(s) && TBG_ALIGNED16(gx),
                "tbg_modulate_bwd: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const RedGeom g = red_geom(B, HW, C / 8, 1, reinterpret_cast<const void*>(&modulate_bwd_kernel));
  modulate_bwd_kernel<<<dim3(g.chunks, B), g.threads, g.smem, stream>>>(
      reinterpret_cast<const uint4*>(gxs), reinterpret_cast<const uint4*>(x), s, reinterpret_cast<uint4*>(gx), gs, HW,
      C / 8, g.pix_per_cta);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_rowdot(const void* a, const void* b, float* out, int B, int HW, int C, void* stream_v) {
  TBG_CHECK_ARG(a && b && out, "tbg_rowdot: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && B >= 1 && HW >= 1, "tbg_rowdot: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(a) && TBG_ALIGNED16(b), "tbg_rowdot: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const RedGeom g = red_geom(B, HW, C / 8, 1, reinterpret_cast<const void*>(&modulate_bwd_kernel));
  modulate_bwd_kernel<<<dim3(g.chunks, B), g.threads, g.smem, stream>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), nullptr, nullptr, out, HW, C / 8, g.pix_per_cta);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_bias_act_fwd(const void* t, const float* noise, const float* noise_strength, const float* bias, void* out,
                                int B, int HW, int C, int act, float gain, void* stream_v) {
  TBG_CHECK_ARG(t && out, "tbg_bias_act_fwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && HW >= 1, "tbg_bias_act_fwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(!noise || noise_strength, "tbg_bias_act_fwd: noise without noise_strength");
  TBG_CHECK_ARG(act >= 0 && act <= 2, "tbg_bias_act_fwd: act must be 0, 1 or 2");
  TBG_CHECK_ARG(TBG_ALIGNED16(t) && TBG_ALIGNED16(out) && TBG_ALIGNED16(bias), "tbg_bias_act_fwd: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n_vec = static_cast<long long>(B) * HW * (C / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = static_cast<long long>(sms()) * 16;
  if (blocks > cap) blocks = cap;
  bias_act_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(t), noise, noise_strength,
                                                                    bias, reinterpret_cast<uint4*>(out), n_vec, HW, C / 8, act,
                                                                    gain);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_bias_act_bwd(const void* g_out, const void* out, const void* residual, const float* noise,
                                const float* d, void* gy0, float* S1, float* Spre, float* Snz, int B, int HW, int C,
                                int act, float gain, int s1_over_batch, void* stream_v) {
  TBG_CHECK_ARG(g_out && out && gy0, "tbg_bias_act_bwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && B >= 1 && HW >= 1, "tbg_bias_act_bwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(!noise || Snz, "tbg_bias_act_bwd: noise without Snz");
  TBG_CHECK_ARG(S1 || !Spre, "tbg_bias_act_bwd: Spre without S1");
  TBG_CHECK_ARG(!s1_over_batch || (!Spre && !noise), "tbg_bias_act_bwd: s1_over_batch yields S1[C] only");
  TBG_CHECK_ARG(TBG_ALIGNED16(g_out) && TBG_ALIGNED16(out) && TBG_ALIGNED16(residual) && TBG_ALIGNED16(d) && TBG_ALIGNED16(gy0),
                "tbg_bias_act_bwd: 16-byte alignment required");
  TBG_CHECK_ARG(gain > 0.f, "tbg_bias_act_bwd: gain must be positive");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int mode = S1 == nullptr ? 0 : ((Spre != nullptr || noise != nullptr) ? 2 : 1);
  static bool attr = false;
  if (!attr) {                                       // before the occupancy query inside red_geom
    cudaFuncSetAttribute(bias_act_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bias_act_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(torgb_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const RedGeom g = red_geom(B, HW, C / 8, mode == 2 ? 3 : 1,
                             mode == 2 ? reinterpret_cast<const void*>(&bias_act_bwd_kernel<2>)
                                       : (mode == 1 ? reinterpret_cast<const void*>(&bias_act_bwd_kernel<1>)
                                                    : reinterpret_cast<const void*>(&bias_act_bwd_kernel<0>)));
  const dim3 grid(g.chunks, B);
#define TBG_BAB(M, SMEM)                                                                                                  \
  bias_act_bwd_kernel<M><<<grid, g.threads, SMEM, stream>>>(                                                              \
      reinterpret_cast<const uint4*>(g_out), reinterpret_cast<const uint4*>(out), reinterpret_cast<const uint4*>(residual), \
      noise, d, reinterpret_cast<uint4*>(gy0), S1, Spre, Snz, HW, C / 8, g.pix_per_cta, act, gain, s1_over_batch)
  if (mode == 0) TBG_BAB(0, 0);
  else if (mode == 1) TBG_BAB(1, g.smem);
  else TBG_BAB(2, g.smem);
#undef TBG_BAB
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_bias_act_rgb_bwd(const void* g_out, const void* out, const float* noise, const float* d,
                                    const float* g_rgb, const float* ws, void* gy0, float* S1, float* Spre, float* Snz,
                                    float* gws, int B, int HW, int C, int act, float gain, void* stream_v) {
  TBG_CHECK_ARG(out && d && g_rgb && ws && gy0 && S1 && Spre && gws, "tbg_bias_act_rgb_bwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && B >= 1 && HW >= 1, "tbg_bias_act_rgb_bwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(!noise || Snz, "tbg_bias_act_rgb_bwd: noise without Snz");
  TBG_CHECK_ARG(act >= 0 && act <= 2 && gain > 0.f, "tbg_bias_act_rgb_bwd: bad activation / gain");
  TBG_CHECK_ARG(TBG_ALIGNED16(g_out) && TBG_ALIGNED16(out) && TBG_ALIGNED16(d) && TBG_ALIGNED16(gy0),
                "tbg_bias_act_rgb_bwd: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(bias_act_rgb_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(bias_act_rgb_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const RedGeom g = red_geom(B, HW, C / 8, 6,
                             g_out != nullptr ? reinterpret_cast<const void*>(&bias_act_rgb_bwd_kernel<true>)
                                              : reinterpret_cast<const void*>(&bias_act_rgb_bwd_kernel<false>));
  TBG_CHECK_ARG(g.smem <= 96 * 1024, "tbg_bias_act_rgb_bwd: C=%d needs too much shared memory", C);
  const dim3 grid(g.chunks, B);
  if (g_out != nullptr)
    bias_act_rgb_bwd_kernel<true><<<grid, g.threads, g.smem, stream>>>(
        reinterpret_cast<const uint4*>(g_out), reinterpret_cast<const uint4*>(out), noise, d, g_rgb, ws,
        reinterpret_cast<uint4*>(gy0), S1, Spre, Snz, gws, HW, C / 8, g.pix_per_cta, act, gain);
  else
    bias_act_rgb_bwd_kernel<false><<<grid, g.threads, g.smem, stream>>>(
        nullptr, reinterpret_cast<const uint4*>(out), noise, d, g_rgb, ws, reinterpret_cast<uint4*>(gy0), S1, Spre, Snz, gws,
        HW, C / 8, g.pix_per_cta, act, gain);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_torgb_fwd(const void* x, const float* ws, const float* bias, float* y, int B, int HW, int C,
                             void* stream_v) {
  TBG_CHECK_ARG(x && ws && y, "tbg_torgb_fwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && HW >= 1, "tbg_torgb_fwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(x), "tbg_torgb_fwd: x must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n_pix = static_cast<long long>(B) * HW;
  long long blocks = (n_pix + 7) / 8;
  const long long cap = static_cast<long long>(sms()) * 8;
  if (blocks > cap) blocks = cap;
  torgb_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(reinterpret_cast<const uint4*>(x), ws, bias, y, B, HW, C / 8);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_torgb_bwd(const void* x, const float* ws, const float* gy, void* gx, float* gws, int B, int HW, int C,
                             void* stream_v) {
  TBG_CHECK_ARG(x && ws && gy && gx && gws, "tbg_torgb_bwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && C <= 2048 && B >= 1 && HW >= 1, "tbg_torgb_bwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(x) && TBG_ALIGNED16(gx), "tbg_torgb_bwd: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(torgb_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const RedGeom g = red_geom(B, HW, C / 8, 3, reinterpret_cast<const void*>(&torgb_bwd_kernel));
  torgb_bwd_kernel<<<dim3(g.chunks, B), g.threads, g.smem, stream>>>(reinterpret_cast<const uint4*>(x), ws, gy,
                                                                      reinterpret_cast<uint4*>(gx), gws, HW, C / 8,
                                                                      g.pix_per_cta);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_fir4(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int offy, int offx,
                        float scale, const float* d, const float* noise, const float* noise_strength, const float* bias,
                        int act, float gain, void* stream_v) {
  TBG_CHECK_ARG(in && out, "tbg_fir4: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && IH >= 1 && IW >= 1 && OH >= 1 && OW >= 1,
                "tbg_fir4: bad shape B=%d in=%dx%d out=%dx%d C=%d", B, IH, IW, OH, OW, C);
  TBG_CHECK_ARG(!noise || noise_strength, "tbg_fir4: noise without noise_strength");
  TBG_CHECK_ARG(act == 0 || act == 1, "tbg_fir4: act must be 0 (linear) or 1 (lrelu)");
  TBG_CHECK_ARG(TBG_ALIGNED16(in) && TBG_ALIGNED16(out) && TBG_ALIGNED16(d), "tbg_fir4: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int cgroups = (C / 8 + 7) / 8;
  TBG_CHECK_ARG(static_cast<long long>(B) * cgroups <= 65535, "tbg_fir4: B * channel groups exceeds the grid limit");
  const dim3 grid((OW + kFirCols - 1) / kFirCols, (OH + kFirStrip - 1) / kFirStrip, B * cgroups);
  const bool epi = d != nullptr || noise != nullptr || bias != nullptr || act != 0 || gain != 1.f;
  if (epi)
    fir4_kernel<true><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), IH, IW, OH,
                                                OW, C / 8, offy, offx, scale, d, noise, noise_strength, bias, act, gain, cgroups);
  else
    fir4_kernel<false><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), IH, IW,
                                                 OH, OW, C / 8, offy, offx, scale, d, noise, noise_strength, bias, act, gain,
                                                 cgroups);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_fir4_down(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int sy, int offy,
                             int offx, float scale, void* stream_v) {
  TBG_CHECK_ARG(in && out, "tbg_fir4_down: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && IH >= 1 && IW >= 1 && OH >= 1 && OW >= 1,
                "tbg_fir4_down: bad shape B=%d in=%dx%d out=%dx%d C=%d", B, IH, IW, OH, OW, C);
  TBG_CHECK_ARG(sy == 1 || sy == 2, "tbg_fir4_down: sy must be 1 or 2 (the width stride is always 2)");
  TBG_CHECK_ARG(TBG_ALIGNED16(in) && TBG_ALIGNED16(out), "tbg_fir4_down: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int cgroups = (C / 8 + 7) / 8;
  TBG_CHECK_ARG(static_cast<long long>(B) * cgroups <= 65535, "tbg_fir4_down: B * channel groups exceeds the grid limit");
  const dim3 grid((OW + kFirCols - 1) / kFirCols, (OH + kFirStrip - 1) / kFirStrip, B * cgroups);
  if (sy == 1)
    fir4_down_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), IH, IW,
                                                  OH, OW, C / 8, offy, offx, scale, cgroups);
  else
    fir4_down_kernel<2><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), IH, IW,
                                                  OH, OW, C / 8, offy, offx, scale, cgroups);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_fir4_down_adjoint(const void* g, const void* add, void* out, int B, int IH, int IW, int OH, int OW, int C,
                                     int sy, int offy, int offx, float scale, void* stream_v) {
  TBG_CHECK_ARG(g && out, "tbg_fir4_down_adjoint: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && IH >= 1 && IW >= 1 && OH >= 1 && OW >= 1,
                "tbg_fir4_down_adjoint: bad shape B=%d in=%dx%d out=%dx%d C=%d", B, IH, IW, OH, OW, C);
  TBG_CHECK_ARG(sy == 1 || sy == 2, "tbg_fir4_down_adjoint: sy must be 1 or 2 (the width stride is always 2)");
  TBG_CHECK_ARG(TBG_ALIGNED16(g) && TBG_ALIGNED16(add) && TBG_ALIGNED16(out), "tbg_fir4_down_adjoint: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int cgroups = (C / 8 + 7) / 8;
  TBG_CHECK_ARG(static_cast<long long>(B) * cgroups <= 65535, "tbg_fir4_down_adjoint: B * channel groups exceeds the grid limit");
  const dim3 grid((IW + kFirCols - 1) / kFirCols, (IH + kFirStrip - 1) / kFirStrip, B * cgroups);
  if (sy == 1)
    fir4_down_adjoint_kernel<1><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(add),
                                                          reinterpret_cast<uint4*>(out), IH, IW, OH, OW, C / 8, offy, offx,
                                                          scale, cgroups);
  else
    fir4_down_adjoint_kernel<2><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(add),
                                                          reinterpret_cast<uint4*>(out), IH, IW, OH, OW, C / 8, offy, offx,
                                                          scale, cgroups);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_fromrgb_fwd(const float* img, const float* w, const float* bias, void* out, int B, int HW, int C,
                               float coef, float gain, void* stream_v) {
  TBG_CHECK_ARG(img && w && bias && out, "tbg_fromrgb_fwd: null pointer");
  TBG_CHECK_ARG(C % 8 == 0 && C >= 8 && B >= 1 && HW >= 1, "tbg_fromrgb_fwd: bad shape B=%d HW=%d C=%d", B, HW, C);
  TBG_CHECK_ARG(TBG_ALIGNED16(out), "tbg_fromrgb_fwd: out must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n_vec = static_cast<long long>(B) * HW * (C / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = static_cast<long long>(sms()) * 16;
  if (blocks > cap) blocks = cap;
  const int c8 = C / 8;
  if (256 % c8 == 0 && B <= 65535 && HW <= (1 << 28)) {
    // whole waves of resident CTAs, each long enough to amortise its 32 weight loads (cost model of red_geom)
    const RedGeom rg = red_geom(B, HW, c8, 0, reinterpret_cast<const void*>(&fromrgb_fwd_cv_kernel));
    const int pix_per_cta = rg.pix_per_cta;
    fromrgb_fwd_cv_kernel<<<dim3((HW + pix_per_cta - 1) / pix_per_cta, B), 256, 0, stream>>>(
        img, w, bias, reinterpret_cast<uint4*>(out), HW, c8, coef, gain, pix_per_cta);
  } else {
    fromrgb_fwd_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(img, w, bias, reinterpret_cast<uint4*>(out), HW, c8,
                                                                     coef, gain, n_vec);
  }
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_fromrgb_bwd(const float* img, const float* w, const void* g_out, const void* out, float* gimg, float* gw,
                               float* gb, int B, int HW, int C, float coef, float gain, void* stream_v) {
  TBG_CHECK_ARG(img && w && g_out && out, "tbg_fromrgb_bwd: null pointer");
  TBG_CHECK_ARG((gw == nullptr) == (gb == nullptr), "tbg_fromrgb_bwd: gw and gb go together");
  const int c8 = C / 8;
  TBG_CHECK_ARG(C % 8 == 0 && c8 >= 1 && c8 <= 32 && (c8 & (c8 - 1)) == 0 && B >= 1 && HW >= 1,
                "tbg_fromrgb_bwd: C=%d must be 8 * a power of two <= 256", C);
  TBG_CHECK_ARG(TBG_ALIGNED16(g_out) && TBG_ALIGNED16(out), "tbg_fromrgb_bwd: 16-byte alignment required");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(fromrgb_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const RedGeom g = red_geom(B, HW, c8, 4, reinterpret_cast<const void*>(&fromrgb_bwd_kernel));
  fromrgb_bwd_kernel<<<dim3(g.chunks, B), g.threads, g.smem, stream>>>(
      img, w, reinterpret_cast<const uint4*>(g_out), reinterpret_cast<const uint4*>(out), gimg, gw, gb, HW, c8, g.pix_per_cta,
      coef, gain, gw != nullptr);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
