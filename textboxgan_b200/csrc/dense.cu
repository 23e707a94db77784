// Small fp32 dense layers of the hot path: Dense.call (dense.py:23-29: flatten -> x @ (coef*W)), the mapping network
// (mapping_block.py:15-45: pixel norm, 5 x [Dense(lrmul .01) + bias*.01 + lrelu*sqrt2]), the word encoder
// (word_encoder.py:39-63: embedding lookup, dropout, Keras Dense(256) + ReLU, reshape/transpose to the base feature map)
// and the discriminator head (discriminator.py:132-142, 213).  All are parameter-sized GEMMs (M = batch or batch*chars,
// K, N <= 8192): exact fp32 FMA on the CUDA cores (the reference computes them in fp32; a TF32 library GEMM would be
// narrower), 64 x 64 output tile per CTA (4 x 4 per thread), K streamed through shared memory in 16-deep slabs, split-K
// over CTAs (fp32 atomics + a finishing pass) when a long reduction meets few output tiles.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

// C(i,j) = alpha * sum_k A(i,k) B(k,j), generic element strides (row- or column-major operands, transposes).
struct GemmOperand {
  const float* p;
  long long s_outer;   // stride of the non-reduction index (i for A, j for B)
  long long s_k;       // stride of the reduction index
};

struct DenseEpilogue {
  const float* bias;   // [N] or null
  float bias_coef;
  int act;             // 0 linear, 1 leaky-relu(0.2), 2 relu
  float gain;
  int accumulate;      // 1: C += result (no bias / activation)
};

static constexpr int kDM = 64, kDN = 64, kDK = 16;   // CTA tile (4 x 4 outputs per thread, 256 threads)

__device__ __forceinline__ float dense_epilogue(float r, int j, const DenseEpilogue& ep) {
  if (ep.bias) r = fmaf(__ldg(ep.bias + j), ep.bias_coef, r);
  if (ep.act == 1) r = r > 0.f ? r : 0.2f * r;
  else if (ep.act == 2) r = fmaxf(r, 0.f);
  return r * ep.gain;
}

// blockIdx.z = K split: split s handles k in [s*k_per_split, (s+1)*k_per_split).  With one split the epilogue is applied
// here; with several the partial products (times alpha) are added atomically into C, which the host entry zeroed, and
// dense_finish_kernel applies the epilogue afterwards.
__global__ void __launch_bounds__(256)
dense_gemm_kernel(GemmOperand A, GemmOperand Bm, float* __restrict__ C, long long ldc, int M, int N, int K, int k_per_split,
                  float alpha, DenseEpilogue ep) {
  __shared__ __align__(16) float As[kDK][kDM + 4];   // [k][i]
  __shared__ __align__(16) float Bs[kDK][kDN + 4];   // [k][j]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads
  const int i0 = blockIdx.y * kDM, j0 = blockIdx.x * kDN;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  // loader mapping: consecutive threads along whichever index is contiguous in memory
  const bool a_kfast = (A.s_k == 1), b_kfast = (Bm.s_k == 1);
  int a_i[4], a_k[4], b_j[4], b_k[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int e = threadIdx.x + 256 * r;            // 1024 elements per operand tile
    a_i[r] = a_kfast ? (e >> 4) : (e & 63);
    a_k[r] = a_kfast ? (e & 15) : (e >> 6);
    b_j[r] = b_kfast ? (e >> 4) : (e & 63);
    b_k[r] = b_kfast ? (e & 15) : (e >> 6);
  }
  float ra[4], rb[4];                                // next slab, prefetched into registers during the FMAs
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int gi = i0 + a_i[r], gka = k0 + a_k[r], gj = j0 + b_j[r], gkb = k0 + b_k[r];
      ra[r] = (gi < M && gka < k_end) ? __ldg(A.p + gi * A.s_outer + gka * A.s_k) : 0.f;
      rb[r] = (gj < N && gkb < k_end) ? __ldg(Bm.p + gj * Bm.s_outer + gkb * Bm.s_k) : 0.f;
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += kDK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      As[a_k[r]][a_i[r]] = ra[r];
      Bs[b_k[r]][b_j[r]] = rb[r];
    }
    __syncthreads();
    if (k0 + kDK < k_end) fetch(k0 + kDK);
#pragma unroll
    for (int k = 0; k < kDK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i >= M || j >= N) continue;
      const float r = acc[u][v] * alpha;
      float* dst = C + i * ldc + j;
      if (split) atomicAdd(dst, r);
      else if (ep.accumulate) *dst += r;
      else *dst = dense_epilogue(r, j, ep);
    }
  }
}

__global__ void dense_finish_kernel(float* __restrict__ C, long long ldc, int M, int N, DenseEpilogue ep) {
  const long long total = static_cast<long long>(M) * N;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(e / N), j = static_cast<int>(e - static_cast<long long>(i) * N);
    float* dst = C + i * ldc + j;
    *dst = dense_epilogue(*dst, j, ep);
  }
}

// gpre[m,n] = gy[m,n] * gain * act'(y[m,n]) (the slope is recovered from the sign of the output: gain > 0);
// gb[n] = bias_coef * sum_m gpre[m,n].  One thread per column, rows streamed (coalesced across the warp).
__global__ void __launch_bounds__(256)
dense_gpre_kernel(const float* __restrict__ y, const float* __restrict__ gy, float* __restrict__ gpre,
                  float* __restrict__ gb, int M, int N, int act, float gain, float bias_coef) {
  // 32 columns x 8 row groups per CTA: coalesced 128-byte rows, the column sums meet in shared memory
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31), tyr = threadIdx.x >> 5;
  float s = 0.f;
  if (n < N) {
#pragma unroll 4
    for (int m = tyr; m < M; m += 8) {
      const size_t o = static_cast<size_t>(m) * N + n;
      float g = gy[o] * gain;
      if (act == 1) g *= (y[o] > 0.f ? 1.f : 0.2f);
      else if (act == 2) g = (y[o] > 0.f ? g : 0.f);
      gpre[o] = g;
      s += g;
    }
  }
  if (!gb) return;
  red[tyr][threadIdx.x & 31] = s;
  __syncthreads();
  if (tyr == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    gb[n] = t * bias_coef;
  }
}

// Pixel norm of the mapping network (mapping_block.py:15-18): y = x * rsqrt(mean_k x^2 + 1e-8); one warp per row.
__global__ void pixel_norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int M, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<size_t>(row) * K;
  float ss = 0.f;
  for (int k = lane; k < K; k += 32) ss = fmaf(xr[k], xr[k], ss);
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = rsqrtf(ss / K + 1e-8f);
  for (int k = lane; k < K; k += 32) y[static_cast<size_t>(row) * K + k] = xr[k] * r;
}
// gx = r * (gy - x * (sum_k gy x) * r^2 / K),  r = rsqrt(mean x^2 + eps)
__global__ void pixel_norm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                      int M, int K) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + static_cast<size_t>(row) * K;
  const float* gr = gy + static_cast<size_t>(row) * K;
  float ss = 0.f, dot = 0.f;
  for (int k = lane; k < K; k += 32) {
    ss = fmaf(xr[k], xr[k], ss);
    dot = fmaf(xr[k], gr[k], dot);
  }
  for (int o = 16; o; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
  }
  const float r = rsqrtf(ss / K + 1e-8f);
  const float c = dot * r * r / K;
  for (int k = lane; k < K; k += 32) gx[static_cast<size_t>(row) * K + k] = r * (gr[k] - xr[k] * c);
}

// ---------------------------------------------------------------------------------------------------------------
// Word encoder (word_encoder.py:39-63).  One CTA per (sample, character) row, one thread per dense output unit.
//   emb = concat(w0[1,E], w[V-1,E])[word] * mask / keep      (Embedding lookups :43-46, Dropout :47)
//   act = relu(emb @ fc_w + fc_b)                            (Keras Dense + ReLU :52-53)
//   out[b, h, w, c] (NHWC bf16) = reshape(act[b], [out_w, out_c, out_h])[w, c, h]   (reshape/transpose :55-63)
// emb [M, E] and act [M, D] (fp32) are saved for the backward pass.
// ---------------------------------------------------------------------------------------------------------------
__global__ void word_encoder_fwd_kernel(const int* __restrict__ words, const float* __restrict__ w0,
                                        const float* __restrict__ table, const float* __restrict__ mask, float inv_keep,
                                        const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                        float* __restrict__ emb_out, float* __restrict__ act_out,
                                        __nv_bfloat16* __restrict__ out, int mcn, int E, int D, int out_h, int out_w,
                                        int out_c) {
  extern __shared__ float s_emb[];
  const int row = blockIdx.x;                 // b * mcn + ch
  const int b = row / mcn, ch = row - b * mcn;
  const int word = words[row];
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    float v = (word == 0) ? w0[i] : table[static_cast<size_t>(word - 1) * E + i];
    if (mask) v *= mask[static_cast<size_t>(row) * E + i] * inv_keep;
    s_emb[i] = v;
    emb_out[static_cast<size_t>(row) * E + i] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float acc = fc_b[j];
    for (int i = 0; i < E; ++i) acc = fmaf(s_emb[i], __ldg(fc_w + static_cast<size_t>(i) * D + j), acc);
    acc = fmaxf(acc, 0.f);
    act_out[static_cast<size_t>(row) * D + j] = acc;
    const int flat = ch * D + j;              // index inside the sample's [out_w, out_c, out_h] block
    const int h = flat % out_h, c = (flat / out_h) % out_c, w = flat / (out_h * out_c);
    out[((static_cast<size_t>(b) * out_h + h) * out_w + w) * out_c + c] = __float2bfloat16(acc);
  }
}

// gpre[row, j] = g_out[b,h,w,c] * (act > 0);   g_table[word-1, i] += (gpre @ fc_w^T)[i] * mask/keep  (row 0 = the frozen
// zero embedding gets no gradient, word_encoder.py:29-33)
__global__ void word_encoder_bwd_kernel(const int* __restrict__ words, const float* __restrict__ mask, float inv_keep,
                                        const float* __restrict__ fc_w, const float* __restrict__ act,
                                        const __nv_bfloat16* __restrict__ g_out, float* __restrict__ gpre,
                                        float* __restrict__ g_table, int mcn, int E, int D, int out_h, int out_w, int out_c) {
  extern __shared__ float s_g[];              // [D]
  const int row = blockIdx.x;
  const int b = row / mcn, ch = row - b * mcn;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    const int flat = ch * D + j;
    const int h = flat % out_h, c = (flat / out_h) % out_c, w = flat / (out_h * out_c);
    float g = __bfloat162float(g_out[((static_cast<size_t>(b) * out_h + h) * out_w + w) * out_c + c]);
    if (!(act[static_cast<size_t>(row) * D + j] > 0.f)) g = 0.f;
    s_g[j] = g;
    gpre[static_cast<size_t>(row) * D + j] = g;
  }
  __syncthreads();
  const int word = words[row];
  if (word == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int i = warp; i < E; i += nwarps) {
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s = fmaf(s_g[j], __ldg(fc_w + static_cast<size_t>(i) * D + j), s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (mask) s *= mask[static_cast<size_t>(row) * E + i] * inv_keep;
      atomicAdd(g_table + static_cast<size_t>(word - 1) * E + i, s);
    }
  }
}

static int launch_gemm(GemmOperand A, GemmOperand Bm, float* C, long long ldc, int M, int N, int K, float alpha,
                       DenseEpilogue ep, cudaStream_t stream) {
  dim3 grid((N + kDN - 1) / kDN, (M + kDM - 1) / kDM, 1);
  // long reductions over few output tiles (the 8192 -> 512 discriminator dense layer at batch 64-128): split K so that
  // about two CTAs per SM are busy, each split at least 128 deep
  int splits = 1;
  const int tiles = grid.x * grid.y;
  if (tiles < 148 && K >= 512) {
    splits = (2 * 148 + tiles - 1) / tiles;
    if (splits > K / 128) splits = K / 128;
    if (splits < 1) splits = 1;
  }
  int k_per_split = ((K + splits - 1) / splits + kDK - 1) / kDK * kDK;
  splits = (K + k_per_split - 1) / k_per_split;
  grid.z = splits;
  if (splits > 1 && !ep.accumulate) {
    // partial products are added atomically: start from zero (dense rows: ldc == N for every caller of this path)
    if (ldc != N) return set_error(TBG_ERR_INVALID_ARG, "dense split-K needs a dense output (ldc == N)");
    TBG_CHECK_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * static_cast<size_t>(M) * N, stream));
  }
  dense_gemm_kernel<<<grid, 256, 0, stream>>>(A, Bm, C, ldc, M, N, K, k_per_split, alpha, ep);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  if (splits > 1 && !ep.accumulate && (ep.bias || ep.act || ep.gain != 1.f)) {
    const long long total = static_cast<long long>(M) * N;
    dense_finish_kernel<<<static_cast<int>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184), 256, 0, stream>>>(C, ldc, M, N, ep);
    count_launch();
    TBG_CHECK_CUDA(cudaGetLastError());
  }
  return TBG_OK;
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_dense_fwd(const float* x, const float* w, const float* bias, float* y, int M, int K, int N, float coef,
                             float bias_coef, int act, float gain, void* stream_v) {
  TBG_CHECK_ARG(x && w && y, "tbg_dense_fwd: null pointer");
  TBG_CHECK_ARG(M >= 1 && K >= 1 && N >= 1, "tbg_dense_fwd: bad shape M=%d K=%d N=%d", M, K, N);
  TBG_CHECK_ARG(act >= 0 && act <= 2, "tbg_dense_fwd: act must be 0, 1 or 2");
  TBG_CHECK_ARG(gain > 0.f, "tbg_dense_fwd: gain must be positive (the backward pass recovers the slope from the output sign)");
  GemmOperand A{x, K, 1}, Bm{w, 1, N};
  DenseEpilogue ep{bias, bias_coef, act, gain, 0};
  return launch_gemm(A, Bm, y, N, M, N, K, coef, ep, reinterpret_cast<cudaStream_t>(stream_v));
}

extern "C" int tbg_dense_bwd(const float* x, const float* w, const float* y, const float* gy, float* gpre, float* gx,
                             float* gw, float* gb, int M, int K, int N, float coef, float bias_coef, int act, float gain,
                             int accumulate_gw, void* stream_v) {
  TBG_CHECK_ARG(w && gy && gpre, "tbg_dense_bwd: null pointer");
  TBG_CHECK_ARG(act == 0 || y, "tbg_dense_bwd: the activation gradient needs the forward output y");
  TBG_CHECK_ARG(!gw || x, "tbg_dense_bwd: the weight gradient needs x");
  TBG_CHECK_ARG(M >= 1 && K >= 1 && N >= 1, "tbg_dense_bwd: bad shape M=%d K=%d N=%d", M, K, N);
  TBG_CHECK_ARG(act >= 0 && act <= 2, "tbg_dense_bwd: act must be 0, 1 or 2");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  dense_gpre_kernel<<<(N + 31) / 32, 256, 0, stream>>>(y, gy, gpre, gb, M, N, act, gain, bias_coef);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  DenseEpilogue ep{nullptr, 0.f, 0, 1.f, 0};
  if (gx) {   // gx[M,K] = coef * gpre[M,N] @ w[K,N]^T
    GemmOperand A{gpre, N, 1}, Bm{w, N, 1};
    int rc = launch_gemm(A, Bm, gx, K, M, K, N, coef, ep, stream);
    if (rc) return rc;
  }
  if (gw) {   // gw[K,N] = coef * x[M,K]^T @ gpre[M,N]
    GemmOperand A{x, 1, K}, Bm{gpre, 1, N};
    ep.accumulate = accumulate_gw;
    int rc = launch_gemm(A, Bm, gw, N, K, N, M, coef, ep, stream);
    if (rc) return rc;
  }
  return TBG_OK;
}

extern "C" int tbg_pixel_norm_fwd(const float* x, float* y, int M, int K, void* stream_v) {
  TBG_CHECK_ARG(x && y && M >= 1 && K >= 1, "tbg_pixel_norm_fwd: bad arguments");
  pixel_norm_fwd_kernel<<<(M + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(x, y, M, K);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_pixel_norm_bwd(const float* x, const float* gy, float* gx, int M, int K, void* stream_v) {
  TBG_CHECK_ARG(x && gy && gx && M >= 1 && K >= 1, "tbg_pixel_norm_bwd: bad arguments");
  pixel_norm_bwd_kernel<<<(M + 3) / 4, 128, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(x, gy, gx, M, K);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_word_encoder_fwd(const int* words, const float* w0, const float* table, const float* mask, float keep,
                                    const float* fc_w, const float* fc_b, float* emb, float* act, void* out, int B, int mcn,
                                    int E, int D, int out_h, int out_w, int out_c, void* stream_v) {
  TBG_CHECK_ARG(words && w0 && table && fc_w && fc_b && emb && act && out, "tbg_word_encoder_fwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && mcn >= 1 && E >= 1 && D >= 1, "tbg_word_encoder_fwd: bad shape");
  TBG_CHECK_ARG(mcn * D == out_h * out_w * out_c, "tbg_word_encoder_fwd: mcn*D=%d must equal out_h*out_w*out_c=%d",
                mcn * D, out_h * out_w * out_c);
  TBG_CHECK_ARG(keep > 0.f && keep <= 1.f, "tbg_word_encoder_fwd: keep probability must be in (0, 1]");
  word_encoder_fwd_kernel<<<B * mcn, 256, E * sizeof(float), reinterpret_cast<cudaStream_t>(stream_v)>>>(
      words, w0, table, mask, 1.f / keep, fc_w, fc_b, emb, act, reinterpret_cast<__nv_bfloat16*>(out), mcn, E, D, out_h,
      out_w, out_c);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_word_encoder_bwd(const int* words, const float* mask, float keep, const float* fc_w, const float* emb,
                                    const float* act, const void* g_out, float* gpre, float* g_table, float* g_fc_w,
                                    float* g_fc_b, int B, int mcn, int E, int D, int out_h, int out_w, int out_c,
                                    void* stream_v) {
  TBG_CHECK_ARG(words && fc_w && emb && act && g_out && gpre && g_table && g_fc_w && g_fc_b,
                "tbg_word_encoder_bwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && mcn >= 1 && E >= 1 && D >= 1 && mcn * D == out_h * out_w * out_c, "tbg_word_encoder_bwd: bad shape");
  TBG_CHECK_ARG(keep > 0.f && keep <= 1.f, "tbg_word_encoder_bwd: keep probability must be in (0, 1]");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int M = B * mcn;
  // g_table is accumulated into with atomics: the caller zeroes it (rows of words that do not occur stay zero)
  word_encoder_bwd_kernel<<<M, 256, D * sizeof(float), stream>>>(words, mask, 1.f / keep, fc_w, act,
                                                                   reinterpret_cast<const __nv_bfloat16*>(g_out), gpre,
                                                                   g_table, mcn, E, D, out_h, out_w, out_c);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  // g_fc_w[E,D] = emb[M,E]^T @ gpre[M,D];  g_fc_b[D] = sum_m gpre
  DenseEpilogue ep{nullptr, 0.f, 0, 1.f, 0};
  GemmOperand A{emb, 1, E}, Bm{gpre, 1, D};
  int rc = launch_gemm(A, Bm, g_fc_w, D, E, D, M, 1.f, ep, stream);
  if (rc) return rc;
  dense_gpre_kernel<<<(D + 31) / 32, 256, 0, stream>>>(nullptr, gpre, gpre, g_fc_b, M, D, 0, 1.f, 1.f);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
