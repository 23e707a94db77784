// Weight preparation for the tensor-core convolutions and its transpose (gradient fold).
//
// The reference multiplies the fp32 HWIO master weight by its equalised-LR coefficient
// (commons.py:4-12, modulated_conv2d.py:71, conv.py:53) and lets cuDNN consume it.  Here every
// convolution geometry wants two bf16 GEMM matrices derived LINEARLY from the master weight:
//
//   fwd[(p,q,o), (t,u,i)] = coef * sum_{kh,kw} Ty[p,t,kh] * Tx[q,u,kw] * w[kh,kw,i,o]
//   adj[(p',q',i), (t',u',o)] = coef * sum_{kh,kw} Ty'[p',t',kh] * Tx'[q',u',kw] * w[kh,kw,i,o]
//
// with tiny per-axis tables: identity for plain convs, the FIR-folded phase table for
// upsample_conv_2d (upfirdn_2d_v2.py:65-103), the FIR-folded 6/4-tap table for conv_downsample_2d
// (upfirdn_2d_v2.py:106-113), and their adjoint re-indexings (conv.py::Axis.adjoint).  One launch
// writes both matrices and, optionally, q[i,o] = coef^2 * sum_{kh,kw} w^2 for the demodulation
// coefficient (modulated_conv2d.py:80-82).  wfold is the exact transpose: it folds the fp32
// gradient of the fwd matrix (+ the gradient of q) back onto the master weight.
//
// Tiling: one CTA per 32(i) x 32(o) tile of the master weight, staged in shared memory so that both
// the o-contiguous reads/writes and the i-contiguous ones are coalesced.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct AxisTable {
  int P, T, K;          // phases, taps per phase, master taps
  float v[2 * 6 * 3];   // [P][T][K]
};

struct WPrepParams {
  AxisTable fy, fx, ay, ax;   // forward and adjoint tables per axis
  int KH, KW, I, O, Ipad, Opad;
  float coef;
};

__device__ __forceinline__ float tab(const AxisTable& t, int p, int tt, int k) { return t.v[(p * t.T + tt) * t.K + k]; }

// cf[c][tap] = Ty[p,t,kh] * Tx[q,u,kw] for combination c = ((p*Px + q)*Ty.T + t)*Tx.T + u, tap = kh*KW + kw;
// built once per CTA in shared memory so that the hot loops index shared memory, not kernel parameters.
__device__ __forceinline__ void build_cf(const AxisTable& ty, const AxisTable& tx, int KH, int KW, float (*cf)[9]) {
  const int ncomb = ty.P * tx.P * ty.T * tx.T;
  for (int e = threadIdx.x; e < ncomb * 9; e += blockDim.x) {
    const int c = e / 9, tap = e - c * 9;
    float v = 0.f;
    if (tap < KH * KW) {
      const int u = c % tx.T, t = (c / tx.T) % ty.T, q = (c / (tx.T * ty.T)) % tx.P, pp = c / (tx.T * ty.T * tx.P);
      v = tab(ty, pp, t, tap / KW) * tab(tx, q, u, tap % KW);
    }
    cf[c][tap] = v;
  }
}

// cf[c][tap] = Ty[p,t,kh] * Tx[q,u,kw] for combination c = ((p*Px + q)*Ty.T + t)*Tx.T + u, tap = kh*KW + kw.
__device__ __forceinline__ void build_cf(const AxisTable& ty, const AxisTable& tx, int KH, int KW, float scale,
                                         float (*cf)[9]) {
  const int ncomb = ty.P * tx.P * ty.T * tx.T;
  for (int e = threadIdx.x; e < ncomb * 9; e += blockDim.x) {
    const int c = e / 9, tap = e - c * 9;
    float v = 0.f;
    if (tap < KH * KW) {
      const int u = c % tx.T, t = (c / tx.T) % ty.T, q = (c / (tx.T * ty.T)) % tx.P, pp = c / (tx.T * ty.T * tx.P);
      v = tab(ty, pp, t, tap / KW) * tab(tx, q, u, tap % KW) * scale;
    }
    cf[c][tap] = v;
  }
}

// One CTA per 32(i) x 32(o) tile of the master weight: the tile is read once, each thread keeps the
// nine taps of its elements in registers and emits every (phase, tap) combination of both matrices.
__global__ void __launch_bounds__(256)
wprep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ adj,
             float* __restrict__ q, const WPrepParams p) {
  __shared__ float sw[9][32][33];  // [kh*KW+kw][i][o]
  __shared__ float cff[36][9], cfa[36][9];
  const int tiles_o = p.Opad / 32;
  const int i0 = (blockIdx.x / tiles_o) * 32, o0 = (blockIdx.x % tiles_o) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
  const int taps = p.KH * p.KW;
  const int TTf = p.fy.T * p.fx.T, ncf = p.fy.P * p.fx.P * TTf;
  const int TTa = p.ay.T * p.ax.T, nca = (adj != nullptr) ? p.ay.P * p.ax.P * TTa : 0;
  build_cf(p.fy, p.fx, p.KH, p.KW, 1.f, cff);
  if (adj != nullptr) build_cf(p.ay, p.ax, p.KH, p.KW, 1.f, cfa);
  for (int tp = 0; tp < 9; ++tp)
    for (int r = ty; r < 32; r += 8) {
      const int i = i0 + r, o = o0 + tx;
      sw[tp][r][tx] = (tp < taps && i < p.I && o < p.O)
                          ? __ldg(w + (static_cast<size_t>(tp) * p.I + i) * p.O + o) * p.coef : 0.f;
    }
  __syncthreads();
  const size_t Kf = static_cast<size_t>(TTf) * p.Ipad, Ka = static_cast<size_t>(TTa) * p.Opad;
  for (int r = ty; r < 32; r += 8) {
    float vf[9], va[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      vf[tp] = sw[tp][tx][r];  // element (i = i0+tx, o = o0+r): lanes over i
      va[tp] = sw[tp][r][tx];  // element (i = i0+r, o = o0+tx): lanes over o
    }
    {  // forward matrix: rows (pq, o), cols (tu, i)
      __nv_bfloat16* dst = fwd + static_cast<size_t>(o0 + r) * Kf + i0 + tx;
      const size_t pq_stride = static_cast<size_t>(p.Opad) * Kf;
      for (int pq = 0, c = 0; pq < p.fy.P * p.fx.P; ++pq, dst += pq_stride) {
#pragma unroll 3
        for (int tu = 0; tu < TTf; ++tu, ++c) {
          float acc = 0.f;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) acc = fmaf(cff[c][tp], vf[tp], acc);
          dst[static_cast<size_t>(tu) * p.Ipad] = __float2bfloat16_rn(acc);
        }
      }
    }
    if (nca > 0) {  // adjoint matrix: rows (pq, i), cols (tu, o)
      __nv_bfloat16* dst = adj + static_cast<size_t>(i0 + r) * Ka + o0 + tx;
      const size_t pq_stride = static_cast<size_t>(p.Ipad) * Ka;
      for (int pq = 0, c = 0; pq < p.ay.P * p.ax.P; ++pq, dst += pq_stride) {
#pragma unroll 3
        for (int tu = 0; tu < TTa; ++tu, ++c) {
          float acc = 0.f;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) acc = fmaf(cfa[c][tp], va[tp], acc);
          dst[static_cast<size_t>(tu) * p.Opad] = __float2bfloat16_rn(acc);
        }
      }
    }
    if (q != nullptr) {
      const int i = i0 + r, o = o0 + tx;
      if (i < p.I && o < p.O) {
        float acc = 0.f;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) acc = fmaf(va[tp], va[tp], acc);
        q[static_cast<size_t>(i) * p.O + o] = acc;
      }
    }
  }
}

// gw[tap,i,o] += coef * sum_c cf[c][tap] * gfwd_c[o,i]  (+ 2*coef^2*w*gq[i,o]).  One CTA per weight tile
// streams the (phase, tap) combinations of the gradient matrix through a 32 x 32 transpose buffer.
__global__ void __launch_bounds__(256)
wfold_kernel(const float* __restrict__ gfwd, const float* __restrict__ gq, const float* __restrict__ w,
             float* __restrict__ gw, const WPrepParams p) {
  __shared__ float sg[2][32][33];  // double-buffered [o][i]
  __shared__ float cff[36][9];
  const int tiles_o = p.Opad / 32;
  const int i0 = (blockIdx.x / tiles_o) * 32, o0 = (blockIdx.x % tiles_o) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int TT = p.fy.T * p.fx.T;
  const size_t Kf = static_cast<size_t>(TT) * p.Ipad;
  const int ncomb = p.fy.P * p.fx.P * TT;
  const int taps = p.KH * p.KW;
  build_cf(p.fy, p.fx, p.KH, p.KW, p.coef, cff);
  float acc[4][9];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) acc[k][tp] = 0.f;
  auto load_tile = [&](int c, int buf) {
    const int pq = c / TT, tu = c - pq * TT;
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // row = o, lanes over i (contiguous)
      const int r = ty + 8 * k;
      sg[buf][r][tx] = __ldg(gfwd + (static_cast<size_t>(pq) * p.Opad + o0 + r) * Kf + static_cast<size_t>(tu) * p.Ipad + i0 + tx);
    }
  };
  load_tile(0, 0);
  __syncthreads();
  for (int c = 0; c < ncomb; ++c) {
    if (c + 1 < ncomb) load_tile(c + 1, (c + 1) & 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // element (i = i0 + ty + 8k, o = o0 + tx)
      const float g = sg[c & 1][tx][ty + 8 * k];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) acc[k][tp] = fmaf(cff[c][tp], g, acc[k][tp]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, o = o0 + tx;
    if (i >= p.I || o >= p.O) continue;
    const float gqv = (gq != nullptr) ? 2.f * p.coef * p.coef * __ldg(gq + static_cast<size_t>(i) * p.O + o) : 0.f;
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      if (tp < taps) {
        const size_t idx = (static_cast<size_t>(tp) * p.I + i) * p.O + o;
        float v = acc[k][tp];
        if (gq != nullptr) v = fmaf(gqv, __ldg(w + idx), v);
        gw[idx] += v;
      }
    }
  }
}

}  // namespace tbg

using namespace tbg;

// tables: 4 blocks of [P, T, K, then P*T*K floats] for fy, fx, ay, ax (ay/ax may have P = 0)
static int load_tables(const float* tables, WPrepParams& p) {
  AxisTable* dst[4] = {&p.fy, &p.fx, &p.ay, &p.ax};
  const float* s = tables;
  for (int a = 0; a < 4; ++a) {
    AxisTable& t = *dst[a];
    t.P = static_cast<int>(s[0]);
    t.T = static_cast<int>(s[1]);
    t.K = static_cast<int>(s[2]);
    if (t.P < 0 || t.P > 2 || t.T < 0 || t.T > 6 || t.K < 0 || t.K > 3) return -1;
    for (int i = 0; i < t.P * t.T * t.K; ++i) t.v[i] = s[3 + i];
    s += 3 + 36;
  }
  return 0;
}

extern "C" int tbg_wprep(const float* w, const float* tables, float coef, int KH, int KW, int I, int O, int Ipad,
                         int Opad, void* fwd, void* adj, float* q, void* stream_v) {
  TBG_CHECK_ARG(w && tables && fwd, "tbg_wprep: null pointer");
  TBG_CHECK_ARG(KH >= 1 && KH <= 3 && KW >= 1 && KW <= 3, "tbg_wprep: master kernel must be at most 3x3");
  TBG_CHECK_ARG(Ipad % 32 == 0 && Opad % 32 == 0 && Ipad >= I && Opad >= O && I >= 1 && O >= 1,
                "tbg_wprep: bad channel counts I=%d O=%d Ipad=%d Opad=%d", I, O, Ipad, Opad);
  WPrepParams p;
  TBG_CHECK_ARG(load_tables(tables, p) == 0, "tbg_wprep: malformed tables");
  TBG_CHECK_ARG(p.fy.K == KH && p.fx.K == KW, "tbg_wprep: tables do not match the kernel size");
  p.KH = KH; p.KW = KW; p.I = I; p.O = O; p.Ipad = Ipad; p.Opad = Opad; p.coef = coef;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const int ncf = p.fy.P * p.fx.P * p.fy.T * p.fx.T;
  const int nca = (p.ay.P > 0) ? p.ay.P * p.ax.P * p.ay.T * p.ax.T : 0;
  TBG_CHECK_ARG(ncf <= 36 && nca <= 36, "tbg_wprep: too many (phase, tap) combinations");
  wprep_kernel<<<(Ipad / 32) * (Opad / 32), 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(fwd),
                                                             (p.ay.P > 0) ? reinterpret_cast<__nv_bfloat16*>(adj) : nullptr,
                                                             q, p);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_wfold(const float* gfwd, const float* gq, const float* w, const float* tables, float coef, int KH,
                         int KW, int I, int O, int Ipad, int Opad, float* gw, void* stream_v) {
  TBG_CHECK_ARG(gfwd && tables && gw, "tbg_wfold: null pointer");
  TBG_CHECK_ARG(!gq || w, "tbg_wfold: gq needs the master weight");
  TBG_CHECK_ARG(KH >= 1 && KH <= 3 && KW >= 1 && KW <= 3, "tbg_wfold: master kernel must be at most 3x3");
  TBG_CHECK_ARG(Ipad % 32 == 0 && Opad % 32 == 0 && Ipad >= I && Opad >= O, "tbg_wfold: bad channel counts");
  WPrepParams p;
  TBG_CHECK_ARG(load_tables(tables, p) == 0, "tbg_wfold: malformed tables");
  TBG_CHECK_ARG(p.fy.P * p.fx.P * p.fy.T * p.fx.T <= 36, "tbg_wfold: too many (phase, tap) combinations");
  p.KH = KH; p.KW = KW; p.I = I; p.O = O; p.Ipad = Ipad; p.Opad = Opad; p.coef = coef;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  wfold_kernel<<<(Ipad / 32) * (Opad / 32), 256, 0, stream>>>(gfwd, gq, w, gw, p);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
