// Parameter-sized algebra of the modulated convolution (modulated_conv2d.py:75-82) and its gradient,
// as single launches instead of chains of library element-wise / reduction / GEMM calls:
//
//   demod_coef   d[b,o]  = rsqrt(sum_i s[b,i]^2 * q[i,o] + eps)                 (:80-82)
//   demod_bwd    from the per-(b,o) sums of bias_act_bwd:
//                  t[b,o]   = dL/d(s^2 @ q) = -0.5 * (Spre - ns*Snz - bias*S1) * d^2
//                  gs[b,i]  = 2 * s[b,i] * sum_o t[b,o] * q[i,o]     (demodulation term of dL/ds;
//                             written, not accumulated: it also initialises the buffer that
//                             modulate_bwd accumulates into)
//                  gbias[o] = sum_b S1[b,o] ;  gns = sum_{b,o} Snz[b,o]
//   (dL/dq[i,o] = sum_b s[b,i]^2 t[b,o] is folded into tbg_wfold.)
//
// All tensors are fp32 and tiny (B x 512, 512 x 512): latency-bound, one wave of CTAs.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

__global__ void __launch_bounds__(256)
demod_coef_kernel(const float* __restrict__ s, const float* __restrict__ q, float* __restrict__ d, int I, int O,
                  float eps) {
  extern __shared__ float sm[];  // s2[I] | red[4][64]
  float* s2 = sm;
  float* red = sm + I;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < I; i += 256) {
    const float v = __ldg(s + static_cast<size_t>(b) * I + i);
    s2[i] = v * v;
  }
  __syncthreads();
  const int ol = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int o = blockIdx.x * 64 + ol;
  float acc = 0.f;
  if (o < O) {
#pragma unroll 4
    for (int i = sl; i < I; i += 4) acc = fmaf(s2[i], __ldg(q + static_cast<size_t>(i) * O + o), acc);
  }
  red[sl * 64 + ol] = acc;
  __syncthreads();
  if (sl == 0 && o < O) d[static_cast<size_t>(b) * O + o] = rsqrtf(red[ol] + red[64 + ol] + red[128 + ol] + red[192 + ol] + eps);
}

__global__ void __launch_bounds__(256)
demod_bwd_kernel(const float* __restrict__ S1, const float* __restrict__ Spre, const float* __restrict__ Snz,
                 const float* __restrict__ d, const float* __restrict__ ns, const float* __restrict__ bias,
                 const float* __restrict__ s, const float* __restrict__ q, float* __restrict__ t_out,
                 float* __restrict__ gbias, float* __restrict__ gns, float* __restrict__ gs, int B, int I, int O,
                 int i_per_cta) {
  extern __shared__ float sm[];  // t[O] (or reduction scratch)
  const int b = blockIdx.y;
  if (b == B) {  // the extra CTA row: reductions over the batch
    for (int o = blockIdx.x * 256 + threadIdx.x; o < O; o += gridDim.x * 256) {
      float acc = 0.f;
      for (int bb = 0; bb < B; ++bb) acc += __ldg(S1 + static_cast<size_t>(bb) * O + o);
      gbias[o] = acc;
    }
    if (blockIdx.x == 0 && gns != nullptr) {
      float acc = 0.f;
      if (Snz != nullptr)
        for (int e = threadIdx.x; e < B * O; e += 256) acc += __ldg(Snz + e);
      sm[threadIdx.x] = acc;
      __syncthreads();
      for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) sm[threadIdx.x] += sm[threadIdx.x + st];
        __syncthreads();
      }
      if (threadIdx.x == 0) gns[0] = sm[0];
    }
    return;
  }
  const float nsv = (ns != nullptr && Snz != nullptr) ? __ldg(ns) : 0.f;
  for (int o = threadIdx.x; o < O; o += 256) {
    const size_t e = static_cast<size_t>(b) * O + o;
    const float dv = __ldg(d + e);
    const float nz = (Snz != nullptr) ? __ldg(Snz + e) : 0.f;
    const float tv = -0.5f * (__ldg(Spre + e) - nsv * nz - __ldg(bias + o) * __ldg(S1 + e)) * dv * dv;
    sm[o] = tv;
    if (blockIdx.x == 0) t_out[e] = tv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * i_per_cta;
  const int i1 = min(I, i0 + i_per_cta);
  for (int i = i0 + warp; i < i1; i += 8) {
    const float* qr = q + static_cast<size_t>(i) * O;
    float acc = 0.f;
    for (int o = lane; o < O; o += 32) acc = fmaf(sm[o], __ldg(qr + o), acc);
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sh);
    if (lane == 0) gs[static_cast<size_t>(b) * I + i] = 2.f * __ldg(s + static_cast<size_t>(b) * I + i) * acc;
  }
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_demod_coef(const float* s, const float* q, float* d, int B, int I, int O, float eps, void* stream_v) {
  TBG_CHECK_ARG(s && q && d, "tbg_demod_coef: null pointer");
  TBG_CHECK_ARG(B >= 1 && I >= 1 && O >= 1 && I <= 8192, "tbg_demod_coef: bad shape B=%d I=%d O=%d", B, I, O);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const size_t smem = (static_cast<size_t>(I) + 256) * sizeof(float);
  demod_coef_kernel<<<dim3((O + 63) / 64, B), 256, smem, stream>>>(s, q, d, I, O, eps);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_demod_bwd(const float* S1, const float* Spre, const float* Snz, const float* d, const float* ns,
                             const float* bias, const float* s, const float* q, float* t, float* gbias, float* gns,
                             float* gs, int B, int I, int O, void* stream_v) {
  TBG_CHECK_ARG(S1 && Spre && d && bias && s && q && t && gbias && gs, "tbg_demod_bwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && I >= 1 && O >= 1 && O <= 8192, "tbg_demod_bwd: bad shape B=%d I=%d O=%d", B, I, O);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int i_per_cta = 64;
  const int chunks = (I + i_per_cta - 1) / i_per_cta;
  const size_t smem = static_cast<size_t>(O > 256 ? O : 256) * sizeof(float);
  demod_bwd_kernel<<<dim3(chunks, B + 1), 256, smem, stream>>>(S1, Spre, Snz, d, ns, bias, s, q, t, gbias, gns, gs, B, I,
                                                              O, i_per_cta);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

// ---------------------------------------------------------------------------------------------
// Grouped style projection: every modulated convolution of the synthesis network maps one row of
// the style tensor through its own dense layer, s_l = mod_bias(mod_dense(style[:, idx_l])) + 1
// (modulated_conv2d.py:75-76, dense.py:23-29).  One launch computes all layers; two launches give
// all gradients (weights + biases, style).
// ---------------------------------------------------------------------------------------------
namespace tbg {

static constexpr int kMaxStyleLayers = 32;

struct StyleLayer {
  const float* w;   // [S, I]
  const float* b;   // [I]
  float* s;         // fwd out [B, I]
  const float* gs;  // bwd in  [B, I]
  float* gw;        // bwd out [S, I]
  float* gb;        // bwd out [I]
  int I, idx;
  int tile0;        // first 32-column tile of this layer in the flattened tile list
};

struct StyleParams {
  StyleLayer l[kMaxStyleLayers];
  int n_layers, total_tiles;
  int B, n_style, S;
  float coef;
};

__device__ __forceinline__ int find_layer(const StyleParams& p, int tile) {
  int l = 0;
  while (l + 1 < p.n_layers && p.l[l + 1].tile0 <= tile) ++l;
  return l;
}

static constexpr int kStRow = 36;   // floats per staged row: 32 batch rows + padding, 16-byte aligned

// grid (total i-tiles, ceil(B/32)); block 256 = 32 i-lanes x 8 K-groups: the reduction axis is split over the
// eight warps (each thread accumulates all 32 batch rows for its share of K, weights read coalesced, the
// staged style rows broadcast as float4), partial sums meet in shared memory.
__global__ void __launch_bounds__(256)
style_dense_fwd_kernel(const __grid_constant__ StyleParams p, const float* __restrict__ style) {
  extern __shared__ __align__(16) float sm_style[];   // st[S][kStRow] | red[8][32][33]
  float* st = sm_style;
  float* red = sm_style + static_cast<size_t>(p.S) * kStRow;
  const int li = find_layer(p, blockIdx.x);
  const StyleLayer& L = p.l[li];
  const int lane = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const int i = (blockIdx.x - L.tile0) * 32 + lane;
  const int b0 = blockIdx.y * 32;
  for (int e = threadIdx.x; e < 32 * p.S; e += 256) {     // lanes over k: contiguous in style
    const int bb = e / p.S, k = e - bb * p.S;
    const int b = b0 + bb;
    st[k * kStRow + bb] = (b < p.B) ? __ldg(style + (static_cast<size_t>(b) * p.n_style + L.idx) * p.S + k) : 0.f;
  }
  __syncthreads();
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  const int kper = (p.S + 7) / 8;
  const int k_begin = kg * kper, k_end = min(p.S, k_begin + kper);
  if (i < L.I) {
#pragma unroll 4
    for (int k = k_begin; k < k_end; ++k) {
      const float wv = __ldg(L.w + static_cast<size_t>(k) * L.I + i);
      const float4* sr = reinterpret_cast<const float4*>(st + k * kStRow);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 sv = sr[q];
        acc[4 * q + 0] = fmaf(sv.x, wv, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(sv.y, wv, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(sv.z, wv, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(sv.w, wv, acc[4 * q + 3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 32; ++r) red[(kg * 32 + r) * 33 + lane] = acc[r];
  __syncthreads();
  if (i < L.I) {
    const float bv = __ldg(L.b + i) + 1.f;
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = kg * 4 + rr, b = b0 + r;
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[(g * 32 + r) * 33 + lane];
      if (b < p.B) L.s[static_cast<size_t>(b) * L.I + i] = fmaf(sum, p.coef, bv);
    }
  }
}

// gw[k,i] = coef * sum_b style[b,idx,k] * gs[b,i] ; gb[i] = sum_b gs[b,i]
// grid (total i-tiles, ceil(S/32)); block 256 = 32 i-lanes x 8 k-groups (4 k rows each)
__global__ void __launch_bounds__(256)
style_dense_wgrad_kernel(const __grid_constant__ StyleParams p, const float* __restrict__ style) {
  __shared__ float st[32][33];  // [k][b chunk]
  const int li = find_layer(p, blockIdx.x);
  const StyleLayer& L = p.l[li];
  const int i = (blockIdx.x - L.tile0) * 32 + (threadIdx.x & 31);
  const int kg = threadIdx.x >> 5;
  const int k0 = blockIdx.y * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float accb = 0.f;
  for (int b0 = 0; b0 < p.B; b0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int bb = e >> 5, kk = e & 31;   // lanes over k: contiguous in style
      const int b = b0 + bb, k = k0 + kk;
      st[kk][bb] = (b < p.B && k < p.S) ? __ldg(style + (static_cast<size_t>(b) * p.n_style + L.idx) * p.S + k) : 0.f;
    }
    __syncthreads();
    if (i < L.I) {
      const int bn = min(32, p.B - b0);
      for (int bb = 0; bb < bn; ++bb) {
        const float g = __ldg(L.gs + static_cast<size_t>(b0 + bb) * L.I + i);
        accb += g;
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r] = fmaf(st[kg * 4 + r][bb], g, acc[r]);
      }
    }
  }
  if (i < L.I) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int k = k0 + kg * 4 + r;
      if (k < p.S) L.gw[static_cast<size_t>(k) * L.I + i] = acc[r] * p.coef;
    }
    if (blockIdx.y == 0 && kg == 0) L.gb[i] = accb;
  }
}

// gstyle[b,j,k] = coef * sum_{l: idx_l == j} sum_i gs_l[b,i] * w_l[k,i]   (zero when no layer uses row j)
// grid (ceil(S/32), n_style, ceil(B/32)); block 256 = 32 k-lanes x 8 i-groups: the reduction axis (the layer's
// input channels) is split over the eight warps, each thread accumulates all 32 batch rows; the layer's
// gradient rows are staged transposed in shared memory and broadcast as float4.
__global__ void __launch_bounds__(256)
style_dense_dgrad_kernel(const __grid_constant__ StyleParams p, float* __restrict__ gstyle, int max_I) {
  extern __shared__ __align__(16) float sm_style[];   // gt[max_I][kStRow] | red[8][32][33]
  float* gt = sm_style;
  float* red = sm_style + static_cast<size_t>(max_I) * kStRow;
  const int k0 = blockIdx.x * 32, j = blockIdx.y, b0 = blockIdx.z * 32;
  const int lane = threadIdx.x & 31, ig = threadIdx.x >> 5;
  const int k = k0 + lane;
  float acc[32];
#pragma unroll
  for (int r = 0; r < 32; ++r) acc[r] = 0.f;
  for (int li = 0; li < p.n_layers; ++li) {
    const StyleLayer& L = p.l[li];
    if (L.idx != j) continue;
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * L.I; e += 256) {   // lanes over i: contiguous in gs
      const int bb = e / L.I, i = e - bb * L.I;
      gt[i * kStRow + bb] = (b0 + bb < p.B) ? __ldg(L.gs + static_cast<size_t>(b0 + bb) * L.I + i) : 0.f;
    }
    __syncthreads();
    const int iper = (((L.I + 7) / 8) + 3) & ~3;         // per-warp share of i, multiple of 4
    const int i_begin = ig * iper, i_end = min(L.I, i_begin + iper);
    if (k < p.S) {
      const float* wr = L.w + static_cast<size_t>(k) * L.I;
      for (int i = i_begin; i < i_end; i += 4) {
        float wv[4];
        if (i + 3 < i_end && ((reinterpret_cast<uintptr_t>(wr + i) & 15) == 0)) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + i));
          wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) wv[u] = (i + u < i_end) ? __ldg(wr + i + u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i + u >= i_end) break;
          const float4* gr = reinterpret_cast<const float4*>(gt + (i + u) * kStRow);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 gv = gr[q];
            acc[4 * q + 0] = fmaf(gv.x, wv[u], acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(gv.y, wv[u], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(gv.z, wv[u], acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(gv.w, wv[u], acc[4 * q + 3]);
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; ++r) red[(ig * 32 + r) * 33 + lane] = acc[r];
  __syncthreads();
  if (k < p.S) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int r = ig * 4 + rr, b = b0 + r;
      float sum = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) sum += red[(g * 32 + r) * 33 + lane];
      if (b < p.B) gstyle[(static_cast<size_t>(b) * p.n_style + j) * p.S + k] = sum * p.coef;
    }
  }
}

}  // namespace tbg

static int fill_style_params(tbg::StyleParams& p, const tbg_style_layer* layers, int n_layers, int B, int n_style, int S,
                             float coef) {
  if (!layers || n_layers < 1 || n_layers > tbg::kMaxStyleLayers) return -1;
  int tiles = 0;
  for (int l = 0; l < n_layers; ++l) {
    if (!layers[l].w || !layers[l].b || layers[l].I < 1 || layers[l].idx < 0 || layers[l].idx >= n_style) return -1;
    p.l[l].w = layers[l].w;
    p.l[l].b = layers[l].b;
    p.l[l].s = layers[l].s;
    p.l[l].gs = layers[l].gs;
    p.l[l].gw = layers[l].gw;
    p.l[l].gb = layers[l].gb;
    p.l[l].I = layers[l].I;
    p.l[l].idx = layers[l].idx;
    p.l[l].tile0 = tiles;
    tiles += (layers[l].I + 31) / 32;
  }
  p.n_layers = n_layers;
  p.total_tiles = tiles;
  p.B = B;
  p.n_style = n_style;
  p.S = S;
  p.coef = coef;
  return 0;
}

extern "C" int tbg_style_dense_fwd(const tbg_style_layer* layers, int n_layers, const float* style, int B, int n_style,
                                   int S, float coef, void* stream_v) {
  TBG_CHECK_ARG(style && B >= 1 && n_style >= 1 && S >= 1, "tbg_style_dense_fwd: bad arguments");
  tbg::StyleParams p;
  TBG_CHECK_ARG(fill_style_params(p, layers, n_layers, B, n_style, S, coef) == 0, "tbg_style_dense_fwd: bad layer table");
  for (int l = 0; l < n_layers; ++l) TBG_CHECK_ARG(layers[l].s, "tbg_style_dense_fwd: null output");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TBG_CHECK_ARG(S <= 1024, "tbg_style_dense_fwd: S=%d exceeds the shared-memory staging (1024)", S);
  const size_t smem = (static_cast<size_t>(S) * tbg::kStRow + 8 * 32 * 33) * sizeof(float);
  static bool attr_f = false;
  if (!attr_f) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(tbg::style_dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_f = true;
  }
  tbg::style_dense_fwd_kernel<<<dim3(p.total_tiles, (B + 31) / 32), 256, smem, stream>>>(p, style);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_style_dense_bwd(const tbg_style_layer* layers, int n_layers, const float* style, float* gstyle, int B,
                                   int n_style, int S, float coef, void* stream_v) {
  TBG_CHECK_ARG(style && gstyle && B >= 1 && n_style >= 1 && S >= 1, "tbg_style_dense_bwd: bad arguments");
  tbg::StyleParams p;
  TBG_CHECK_ARG(fill_style_params(p, layers, n_layers, B, n_style, S, coef) == 0, "tbg_style_dense_bwd: bad layer table");
  for (int l = 0; l < n_layers; ++l)
    TBG_CHECK_ARG(layers[l].gs && layers[l].gw && layers[l].gb, "tbg_style_dense_bwd: null gradient pointer");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  tbg::style_dense_wgrad_kernel<<<dim3(p.total_tiles, (S + 31) / 32), 256, 0, stream>>>(p, style);
  count_launch();
  int max_I = 1;
  for (int l = 0; l < n_layers; ++l) max_I = layers[l].I > max_I ? layers[l].I : max_I;
  TBG_CHECK_ARG(max_I <= 1024, "tbg_style_dense_bwd: layer width %d exceeds the shared-memory staging (1024)", max_I);
  const size_t smem = (static_cast<size_t>(max_I) * tbg::kStRow + 8 * 32 * 33) * sizeof(float);
  static bool attr_d = false;
  if (!attr_d) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(tbg::style_dense_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_d = true;
  }
  tbg::style_dense_dgrad_kernel<<<dim3((S + 31) / 32, n_style, (B + 31) / 32), 256, smem, stream>>>(p, gstyle, max_I);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
