// Whole-sequence LSTM recurrence (forward + input-gradient BPTT) for the frozen BiLSTM stack of the
// ASTER OCR head (aster_ocr_utils/aster_inferer.py:28-37 calls it through the SavedModel; the
// head is frozen, so only d/d(input projection) is needed, never weight gradients).
//
// One launch walks all T steps: grid = directions x ceil(B / BS) CTAs, 256 threads = one per hidden
// unit (H = 256).  The input projections x@W_ih + b for all steps are a plain batched GEMM done
// beforehand; here each step is gates = xp[t] + h @ W_hh (a 256 x 1024 mat-vec per sample, bf16
// weights streamed from L2, fp32 accumulate), the cell update, and h written to shared memory for
// the next step.  Latency-bound by design: the recurrence is sequential in t.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace cg = cooperative_groups;

namespace tbg {

static constexpr int kH = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

static constexpr int kQ = 4;            // the 256-long dot products are split over 4 thread groups
static constexpr int kThreads = kH * kQ;

// w_packed: [D][H (k)][H (j)][4 gates] bf16  (gate order i, f, g, o)
// 1024 threads: thread (q, j) accumulates k in [64q, 64q+64) for hidden unit j — four times the
// loads in flight per SM of a 256-thread version, which is what bounds this latency-bound stream.
template <int BS>
__global__ void __launch_bounds__(kThreads) lstm_fwd_kernel(const float* __restrict__ xp,
                                                            const __nv_bfloat16* __restrict__ w_packed,
                                                            float* __restrict__ h_out, float* __restrict__ gates_out,
                                                            float* __restrict__ c_out, int B, int T, int groups) {
  __shared__ float h_s[BS][kH];
  __shared__ float part[kQ - 1][BS][4][kH];
  const int j = threadIdx.x & (kH - 1);
  const int q = threadIdx.x >> 8;
  const int d = blockIdx.x / groups;
  const int b0 = (blockIdx.x % groups) * BS;
  const uint2* w = reinterpret_cast<const uint2*>(w_packed) + static_cast<size_t>(d) * kH * kH;
  float c[BS];
#pragma unroll
  for (int s = 0; s < BS; ++s) c[s] = 0.f;
  if (q == 0) {
#pragma unroll
    for (int s = 0; s < BS; ++s) h_s[s][j] = 0.f;
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float acc[BS][4];
#pragma unroll
    for (int s = 0; s < BS; ++s) {
      const int b = b0 + s;
      if (q == 0 && b < B) {
        const float* xr = xp + ((static_cast<size_t>(d) * B + b) * T + t) * (4 * kH);
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[s][g] = __ldg(xr + g * kH + j);
      } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[s][g] = 0.f;
      }
    }
    const int kbeg = q * (kH / kQ);
#pragma unroll
    for (int k0 = 0; k0 < kH / kQ; k0 += 16) {
      uint2 wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = __ldg(w + static_cast<size_t>(kbeg + k0 + u) * kH + j);
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[u].x));
        const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[u].y));
#pragma unroll
        for (int s = 0; s < BS; ++s) {
          const float hk = h_s[s][kbeg + k0 + u];
          acc[s][0] = fmaf(hk, w01.x, acc[s][0]);
          acc[s][1] = fmaf(hk, w01.y, acc[s][1]);
          acc[s][2] = fmaf(hk, w23.x, acc[s][2]);
          acc[s][3] = fmaf(hk, w23.y, acc[s][3]);
        }
      }
    }
    if (q > 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s)
#pragma unroll
        for (int g = 0; g < 4; ++g) part[q - 1][s][g][j] = acc[s][g];
    }
    __syncthreads();  // partial sums visible; every thread has finished reading h_s of step t-1
    if (q == 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        const int b = b0 + s;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int qq = 0; qq < kQ - 1; ++qq) acc[s][g] += part[qq][s][g][j];
        const float ig = sigmoidf_(acc[s][0]);
        const float fg = sigmoidf_(acc[s][1]);
        const float gg = tanhf(acc[s][2]);
        const float og = sigmoidf_(acc[s][3]);
        c[s] = fg * c[s] + ig * gg;
        const float h = og * tanhf(c[s]);
        h_s[s][j] = h;
        if (b < B) {
          const size_t row = (static_cast<size_t>(d) * B + b) * T + t;
          float* gr = gates_out + row * (4 * kH);
          gr[j] = ig;
          gr[kH + j] = fg;
          gr[2 * kH + j] = gg;
          gr[3 * kH + j] = og;
          c_out[row * kH + j] = c[s];
          h_out[row * kH + j] = h;
        }
      }
    }
    __syncthreads();
  }
}

// wT_packed: [D][H (j)][H (k)][4 gates] bf16 — the same weights indexed for dh_prev[k] = sum_j,g
template <int BS>
__global__ void __launch_bounds__(kThreads) lstm_bwd_kernel(const float* __restrict__ g_h, const float* __restrict__ gates,
                                                            const float* __restrict__ c_saved,
                                                            const __nv_bfloat16* __restrict__ wT_packed,
                                                            float* __restrict__ g_xp, int B, int T, int groups) {
  __shared__ float dg_s[BS][4][kH];
  __shared__ float part[kQ - 1][BS][kH];
  const int j = threadIdx.x & (kH - 1);
  const int q = threadIdx.x >> 8;
  const int d = blockIdx.x / groups;
  const int b0 = (blockIdx.x % groups) * BS;
  const uint2* w = reinterpret_cast<const uint2*>(wT_packed) + static_cast<size_t>(d) * kH * kH;
  float dh_rec[BS], dc_next[BS];
#pragma unroll
  for (int s = 0; s < BS; ++s) dh_rec[s] = dc_next[s] = 0.f;
  for (int t = T - 1; t >= 0; --t) {
    if (q == 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        const int b = b0 + s;
        float di = 0.f, df = 0.f, dgg = 0.f, dog = 0.f;
        if (b < B) {
          const size_t row = (static_cast<size_t>(d) * B + b) * T + t;
          const float* gr = gates + row * (4 * kH);
          const float ig = __ldg(gr + j), fg = __ldg(gr + kH + j), gg = __ldg(gr + 2 * kH + j), og = __ldg(gr + 3 * kH + j);
          const float cc = __ldg(c_saved + row * kH + j);
          const float cp = (t > 0) ? __ldg(c_saved + (row - 1) * kH + j) : 0.f;
          const float dh = __ldg(g_h + row * kH + j) + dh_rec[s];
          const float tc = tanhf(cc);
          dog = dh * tc * og * (1.f - og);
          const float dc = dh * og * (1.f - tc * tc) + dc_next[s];
          di = dc * gg * ig * (1.f - ig);
          df = dc * cp * fg * (1.f - fg);
          dgg = dc * ig * (1.f - gg * gg);
          dc_next[s] = dc * fg;
          float* go = g_xp + row * (4 * kH);
          go[j] = di;
          go[kH + j] = df;
          go[2 * kH + j] = dgg;
          go[3 * kH + j] = dog;
        }
        dg_s[s][0][j] = di;
        dg_s[s][1][j] = df;
        dg_s[s][2][j] = dgg;
        dg_s[s][3][j] = dog;
      }
    }
    __syncthreads();
    float acc[BS];
#pragma unroll
    for (int s = 0; s < BS; ++s) acc[s] = 0.f;
    const int jbeg = q * (kH / kQ);
#pragma unroll
    for (int j0 = 0; j0 < kH / kQ; j0 += 16) {
      uint2 wv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) wv[u] = __ldg(w + static_cast<size_t>(jbeg + j0 + u) * kH + j);  // thread index plays k
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[u].x));
        const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[u].y));
#pragma unroll
        for (int s = 0; s < BS; ++s) {
          acc[s] = fmaf(dg_s[s][0][jbeg + j0 + u], w01.x, acc[s]);
          acc[s] = fmaf(dg_s[s][1][jbeg + j0 + u], w01.y, acc[s]);
          acc[s] = fmaf(dg_s[s][2][jbeg + j0 + u], w23.x, acc[s]);
          acc[s] = fmaf(dg_s[s][3][jbeg + j0 + u], w23.y, acc[s]);
        }
      }
    }
    if (q > 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) part[q - 1][s][j] = acc[s];
    }
    __syncthreads();
    if (q == 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        float v = acc[s];
#pragma unroll
        for (int qq = 0; qq < kQ - 1; ++qq) v += part[qq][s][j];
        dh_rec[s] = v;
      }
    }
    // (the next iteration's first __syncthreads orders the reads of `part` / `dg_s` before their reuse)
  }
}

// ---------------------------------------------------------------------------------------------------------
// Cluster variants (UNVALIDATED, enabled with TBG_LSTM_CLUSTER=1): the kernels above stream the whole bf16 W_hh
// (512 KB) from L2 every step — 79 B/clk/SM, the per-SM L2 feed limit — on D*B/2 SMs only.  Here a cluster of 4
// CTAs owns BS samples of one direction; CTA r keeps the 128 KB slice of W_hh for hidden units [64r, 64r+64) in
// shared memory for the whole sequence, and the CTAs exchange the 256-wide state once per step through distributed
// shared memory (double-buffered, one cluster barrier per step).
// ---------------------------------------------------------------------------------------------------------
static constexpr int kCl = 4;                 // CTAs per cluster
static constexpr int kUl = kH / kCl;          // hidden units per CTA (64)
static constexpr int kKg = kThreads / kUl;    // reduction groups per CTA (16)
static constexpr int kKper = kH / kKg;        // reduction indices per group (16)

template <int BS>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads)
lstm_fwd_cluster_kernel(const float* __restrict__ xp, const __nv_bfloat16* __restrict__ w_packed,
                        float* __restrict__ h_out, float* __restrict__ gates_out, float* __restrict__ c_out, int B, int T,
                        int groups) {
  extern __shared__ __align__(16) uint8_t sm_lstm[];
  uint2* wsm = reinterpret_cast<uint2*>(sm_lstm);                              // [kH k][kUl units] x 4 gates bf16
  float* h_s = reinterpret_cast<float*>(sm_lstm + kH * kUl * sizeof(uint2));  // [2][BS][kH]
  float* part = h_s + 2 * BS * kH;                                            // [kKg-1][BS][4][kUl]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cl = blockIdx.x / kCl;
  const int d = cl / groups, b0 = (cl % groups) * BS;
  const int ul = threadIdx.x & (kUl - 1), kg = threadIdx.x / kUl;
  const int j = rank * kUl + ul;
  const uint2* w = reinterpret_cast<const uint2*>(w_packed) + static_cast<size_t>(d) * kH * kH;
  for (int e = threadIdx.x; e < kH * kUl; e += kThreads) wsm[e] = __ldg(w + static_cast<size_t>(e / kUl) * kH + rank * kUl + (e % kUl));
  for (int e = threadIdx.x; e < 2 * BS * kH; e += kThreads) h_s[e] = 0.f;
  float c[BS], xn[BS][4];
#pragma unroll
  for (int s = 0; s < BS; ++s) {
    c[s] = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      xn[s][g] = (kg == 0 && b0 + s < B) ? __ldg(xp + ((static_cast<size_t>(d) * B + b0 + s) * T) * (4 * kH) + g * kH + j) : 0.f;
  }
  cluster.sync();
  int cur = 0;
  for (int t = 0; t < T; ++t) {
    float acc[BS][4];
#pragma unroll
    for (int s = 0; s < BS; ++s)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[s][g] = xn[s][g];      // zero for kg > 0
    if (kg == 0 && t + 1 < T) {                               // prefetch the next step's input projection
#pragma unroll
      for (int s = 0; s < BS; ++s)
        if (b0 + s < B) {
          const float* xr = xp + ((static_cast<size_t>(d) * B + b0 + s) * T + t + 1) * (4 * kH);
#pragma unroll
          for (int g = 0; g < 4; ++g) xn[s][g] = __ldg(xr + g * kH + j);
        }
    }
    const float* hc = h_s + cur * BS * kH;
#pragma unroll
    for (int u = 0; u < kKper; ++u) {
      const int k = kg * kKper + u;
      const uint2 wv = wsm[k * kUl + ul];
      const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv.x));
      const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv.y));
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        const float hk = hc[s * kH + k];
        acc[s][0] = fmaf(hk, w01.x, acc[s][0]);
        acc[s][1] = fmaf(hk, w01.y, acc[s][1]);
        acc[s][2] = fmaf(hk, w23.x, acc[s][2]);
        acc[s][3] = fmaf(hk, w23.y, acc[s][3]);
      }
    }
    if (kg > 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s)
#pragma unroll
        for (int g = 0; g < 4; ++g) part[(((kg - 1) * BS + s) * 4 + g) * kUl + ul] = acc[s][g];
    }
    __syncthreads();
    if (kg == 0) {
      const int nxt = cur ^ 1;
#pragma unroll
      for (int s = 0; s < BS; ++s) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          for (int q = 0; q < kKg - 1; ++q) acc[s][g] += part[((q * BS + s) * 4 + g) * kUl + ul];
        const float ig = sigmoidf_(acc[s][0]);
        const float fg = sigmoidf_(acc[s][1]);
        const float gg = tanhf(acc[s][2]);
        const float og = sigmoidf_(acc[s][3]);
        c[s] = fg * c[s] + ig * gg;
        const float h = og * tanhf(c[s]);
#pragma unroll
        for (int r = 0; r < kCl; ++r) cluster.map_shared_rank(h_s, r)[(nxt * BS + s) * kH + j] = h;
        const int b = b0 + s;
        if (b < B) {
          const size_t row = (static_cast<size_t>(d) * B + b) * T + t;
          float* gr = gates_out + row * (4 * kH);
          gr[j] = ig;
          gr[kH + j] = fg;
          gr[2 * kH + j] = gg;
          gr[3 * kH + j] = og;
          c_out[row * kH + j] = c[s];
          h_out[row * kH + j] = h;
        }
      }
    }
    cluster.sync();        // every CTA's slice of h(t) has landed in every CTA; `part` and h_s[cur] may be reused
    cur ^= 1;
  }
}

template <int BS>
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kThreads)
lstm_bwd_cluster_kernel(const float* __restrict__ g_h, const float* __restrict__ gates, const float* __restrict__ c_saved,
                        const __nv_bfloat16* __restrict__ wT_packed, float* __restrict__ g_xp, int B, int T, int groups) {
  extern __shared__ __align__(16) uint8_t sm_lstm[];
  uint2* wsm = reinterpret_cast<uint2*>(sm_lstm);                              // [kH j][kUl units k] x 4 gates bf16
  float* dg_s = reinterpret_cast<float*>(sm_lstm + kH * kUl * sizeof(uint2)); // [2][BS][4][kH]
  float* part = dg_s + 2 * BS * 4 * kH;                                        // [kKg-1][BS][kUl]
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cl = blockIdx.x / kCl;
  const int d = cl / groups, b0 = (cl % groups) * BS;
  const int ul = threadIdx.x & (kUl - 1), kg = threadIdx.x / kUl;
  const int j = rank * kUl + ul;                                               // the unit this thread owns (state index)
  const uint2* w = reinterpret_cast<const uint2*>(wT_packed) + static_cast<size_t>(d) * kH * kH;
  for (int e = threadIdx.x; e < kH * kUl; e += kThreads) wsm[e] = __ldg(w + static_cast<size_t>(e / kUl) * kH + rank * kUl + (e % kUl));
  float dh_rec[BS], dc_next[BS];
#pragma unroll
  for (int s = 0; s < BS; ++s) dh_rec[s] = dc_next[s] = 0.f;
  cluster.sync();
  for (int t = T - 1; t >= 0; --t) {
    const int buf = t & 1;
    if (kg == 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        const int b = b0 + s;
        float di = 0.f, df = 0.f, dgg = 0.f, dog = 0.f;
        if (b < B) {
          const size_t row = (static_cast<size_t>(d) * B + b) * T + t;
          const float* gr = gates + row * (4 * kH);
          const float ig = __ldg(gr + j), fg = __ldg(gr + kH + j), gg = __ldg(gr + 2 * kH + j), og = __ldg(gr + 3 * kH + j);
          const float cc = __ldg(c_saved + row * kH + j);
          const float cp = (t > 0) ? __ldg(c_saved + (row - 1) * kH + j) : 0.f;
          const float dh = __ldg(g_h + row * kH + j) + dh_rec[s];
          const float tc = tanhf(cc);
          dog = dh * tc * og * (1.f - og);
          const float dc = dh * og * (1.f - tc * tc) + dc_next[s];
          di = dc * gg * ig * (1.f - ig);
          df = dc * cp * fg * (1.f - fg);
          dgg = dc * ig * (1.f - gg * gg);
          dc_next[s] = dc * fg;
          float* go = g_xp + row * (4 * kH);
          go[j] = di;
          go[kH + j] = df;
          go[2 * kH + j] = dgg;
          go[3 * kH + j] = dog;
        }
#pragma unroll
        for (int r = 0; r < kCl; ++r) {
          float* rem = cluster.map_shared_rank(dg_s, r) + ((buf * BS + s) * 4) * kH + j;
          rem[0] = di;
          rem[kH] = df;
          rem[2 * kH] = dgg;
          rem[3 * kH] = dog;
        }
      }
    }
    cluster.sync();        // all 256 gate gradients of step t are in every CTA
    float acc[BS];
#pragma unroll
    for (int s = 0; s < BS; ++s) acc[s] = 0.f;
    const float* dgc = dg_s + buf * BS * 4 * kH;
#pragma unroll
    for (int u = 0; u < kKper; ++u) {
      const int jj = kg * kKper + u;
      const uint2 wv = wsm[jj * kUl + ul];
      const float2 w01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv.x));
      const float2 w23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv.y));
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        const float* dgp = dgc + s * 4 * kH + jj;
        acc[s] = fmaf(dgp[0], w01.x, acc[s]);
        acc[s] = fmaf(dgp[kH], w01.y, acc[s]);
        acc[s] = fmaf(dgp[2 * kH], w23.x, acc[s]);
        acc[s] = fmaf(dgp[3 * kH], w23.y, acc[s]);
      }
    }
    if (kg > 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) part[((kg - 1) * BS + s) * kUl + ul] = acc[s];
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
      for (int s = 0; s < BS; ++s) {
        float v = acc[s];
        for (int q = 0; q < kKg - 1; ++q) v += part[(q * BS + s) * kUl + ul];
        dh_rec[s] = v;
      }
    }
    __syncthreads();       // `part` is rewritten in the next step
  }
}

static bool lstm_cluster_enabled() { return g_tuning.lstm_cluster != 0; }

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_lstm_seq_fwd(const float* xp, const void* w_packed, float* h_out, float* gates_out, float* c_out,
                                int D, int B, int T, int H, void* stream_v) {
  TBG_CHECK_ARG(xp && w_packed && h_out && gates_out && c_out, "tbg_lstm_seq_fwd: null pointer");
  TBG_CHECK_ARG(H == kH, "tbg_lstm_seq_fwd: hidden size must be %d (got %d)", kH, H);
  TBG_CHECK_ARG(D >= 1 && B >= 1 && T >= 1, "tbg_lstm_seq_fwd: bad shape D=%d B=%d T=%d", D, B, T);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (lstm_cluster_enabled() && B >= 4) {
    constexpr int BS = 4;
    const int groups = (B + BS - 1) / BS;
    const size_t smem = kH * kUl * sizeof(uint2) + (2 * BS * kH + (kKg - 1) * BS * 4 * kUl) * sizeof(float);
    static bool attr = false;
    if (!attr) {
      TBG_CHECK_CUDA(cudaFuncSetAttribute(lstm_fwd_cluster_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    lstm_fwd_cluster_kernel<BS><<<D * groups * kCl, kThreads, smem, stream>>>(
        xp, reinterpret_cast<const __nv_bfloat16*>(w_packed), h_out, gates_out, c_out, B, T, groups);
    count_launch();
    TBG_CHECK_CUDA(cudaGetLastError());
    return TBG_OK;
  }
  if (B >= 16) {
    const int groups = (B + 1) / 2;
    lstm_fwd_kernel<2><<<D * groups, kThreads, 0, stream>>>(xp, reinterpret_cast<const __nv_bfloat16*>(w_packed), h_out,
                                                      gates_out, c_out, B, T, groups);
  } else {
    lstm_fwd_kernel<1><<<D * B, kThreads, 0, stream>>>(xp, reinterpret_cast<const __nv_bfloat16*>(w_packed), h_out,
                                                 gates_out, c_out, B, T, B);
  }
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_lstm_seq_bwd(const float* g_h, const float* gates, const float* c_saved, const void* wT_packed,
                                float* g_xp, int D, int B, int T, int H, void* stream_v) {
  TBG_CHECK_ARG(g_h && gates && c_saved && wT_packed && g_xp, "tbg_lstm_seq_bwd: null pointer");
  TBG_CHECK_ARG(H == kH, "tbg_lstm_seq_bwd: hidden size must be %d (got %d)", kH, H);
  TBG_CHECK_ARG(D >= 1 && B >= 1 && T >= 1, "tbg_lstm_seq_bwd: bad shape D=%d B=%d T=%d", D, B, T);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  if (lstm_cluster_enabled() && B >= 4) {
    constexpr int BS = 4;
    const int groups = (B + BS - 1) / BS;
    const size_t smem = kH * kUl * sizeof(uint2) + (2 * BS * 4 * kH + (kKg - 1) * BS * kUl) * sizeof(float);
    static bool attr = false;
    if (!attr) {
      TBG_CHECK_CUDA(cudaFuncSetAttribute(lstm_bwd_cluster_kernel<BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    lstm_bwd_cluster_kernel<BS><<<D * groups * kCl, kThreads, smem, stream>>>(
        g_h, gates, c_saved, reinterpret_cast<const __nv_bfloat16*>(wT_packed), g_xp, B, T, groups);
    count_launch();
    TBG_CHECK_CUDA(cudaGetLastError());
    return TBG_OK;
  }
  if (B >= 16) {
    const int groups = (B + 1) / 2;
    lstm_bwd_kernel<2><<<D * groups, kThreads, 0, stream>>>(g_h, gates, c_saved, reinterpret_cast<const __nv_bfloat16*>(wT_packed),
                                                      g_xp, B, T, groups);
  } else {
    lstm_bwd_kernel<1><<<D * B, kThreads, 0, stream>>>(g_h, gates, c_saved, reinterpret_cast<const __nv_bfloat16*>(wT_packed),
                                                 g_xp, B, T, B);
  }
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
