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
#include <cstring>

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

// cf[c][tap] = Ty[p,t,kh] * Tx[q,u,kw] for combination c = ((p*Px + q)*Ty.T + t)*Tx.T + u, tap = kh*KW + kw.
__device__ __forceinline__ void build_cf(const AxisTable& ty, const AxisTable& tx, int KH, int KW, float scale,
                                         float (*cf)[12]) {
  // rows padded to 12 floats (48 B) so that a combination's nine coefficients are three 16-byte loads
  const int ncomb = ty.P * tx.P * ty.T * tx.T;
  for (int e = threadIdx.x; e < ncomb * 12; e += blockDim.x) {
    const int c = e / 12, tap = e - c * 12;
    float v = 0.f;
    if (tap < KH * KW) {
      const int u = c % tx.T, t = (c / tx.T) % ty.T, q = (c / (tx.T * ty.T)) % tx.P, pp = c / (tx.T * ty.T * tx.P);
      v = tab(ty, pp, t, tap / KW) * tab(tx, q, u, tap % KW) * scale;
    }
    cf[c][tap] = v;
  }
}

// One CTA per 32(i) x 32(o) tile of the master weight and per CHUNK of (phase, tap) combinations
// (blockIdx.y: forward chunks first, then adjoint chunks) so that even a 128 x 128 weight fills the
// SMs; the tile is re-read by each chunk (L2 hits).  A thread owns two neighbouring elements along the
// contiguous axis of the output matrix, keeps their nine taps in registers and emits one packed
// bf16x2 store per combination; per-combination offsets and coefficients come from shared memory.
__device__ __forceinline__ void wprep_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                           __nv_bfloat16* __restrict__ adj, float* __restrict__ q, const WPrepParams& p,
                                           const int chunk, const int fwd_chunks, const int bx, const int by) {
  __shared__ float sw[9][32][33];  // [kh*KW+kw][i][o]
  __shared__ __align__(16) float cf[36][12];
  __shared__ unsigned coff[36];    // element offset of combination c inside the output matrix
  const int tiles_o = p.Opad / 32;
  const int i0 = (bx / tiles_o) * 32, o0 = (bx % tiles_o) * 32;
  const int taps = p.KH * p.KW;
  const bool is_adj = by >= fwd_chunks;
  const int chunk_id = is_adj ? by - fwd_chunks : by;
  const AxisTable& ay = is_adj ? p.ay : p.fy;
  const AxisTable& ax = is_adj ? p.ax : p.fx;
  const int TT = ay.T * ax.T, ncomb = ay.P * ax.P * TT;
  const int c_begin = chunk_id * chunk, c_end = min(ncomb, c_begin + chunk);
  // forward matrix: rows (pq, o), cols (tu, i), contiguous over i;  adjoint: rows (pq, i), cols (tu, o), contiguous over o
  const int rows_pad = is_adj ? p.Ipad : p.Opad, cols_pad = is_adj ? p.Opad : p.Ipad;
  const unsigned Kc = static_cast<unsigned>(TT) * cols_pad;
  build_cf(ay, ax, p.KH, p.KW, 1.f, cf);
  if (threadIdx.x < ncomb) {
    const int c = threadIdx.x, pq = c / TT, tu = c - pq * TT;
    coff[c] = static_cast<unsigned>(pq) * rows_pad * Kc + static_cast<unsigned>(tu) * cols_pad;
  }
  {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int tp = 0; tp < 9; ++tp)
      for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r, o = o0 + tx;
        sw[tp][r][tx] = (tp < taps && i < p.I && o < p.O)
                            ? __ldg(w + (static_cast<size_t>(tp) * p.I + i) * p.O + o) * p.coef : 0.f;
      }
  }
  __syncthreads();
  __nv_bfloat16* const base = is_adj ? adj : fwd;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // 16 column pairs x 16 rows, two row passes
#pragma unroll
  for (int rp = 0; rp < 2; ++rp) {
    const int r = ty + 16 * rp;
    float v0[9], v1[9];
#pragma unroll
    for (int tp = 0; tp < 9; ++tp) {
      v0[tp] = is_adj ? sw[tp][r][2 * tx] : sw[tp][2 * tx][r];
      v1[tp] = is_adj ? sw[tp][r][2 * tx + 1] : sw[tp][2 * tx + 1][r];
    }
    const int row = (is_adj ? i0 : o0) + r, colv = (is_adj ? o0 : i0) + 2 * tx;
    __nv_bfloat16* const dst = base + static_cast<size_t>(row) * Kc + colv;
#pragma unroll 2
    for (int c = c_begin; c < c_end; ++c) {
      const float4 c0 = *reinterpret_cast<const float4*>(&cf[c][0]);
      const float4 c1 = *reinterpret_cast<const float4*>(&cf[c][4]);
      const float c2 = cf[c][8];
      float a0 = c0.x * v0[0], a1 = c0.x * v1[0];
      a0 = fmaf(c0.y, v0[1], a0); a1 = fmaf(c0.y, v1[1], a1);
      a0 = fmaf(c0.z, v0[2], a0); a1 = fmaf(c0.z, v1[2], a1);
      a0 = fmaf(c0.w, v0[3], a0); a1 = fmaf(c0.w, v1[3], a1);
      a0 = fmaf(c1.x, v0[4], a0); a1 = fmaf(c1.x, v1[4], a1);
      a0 = fmaf(c1.y, v0[5], a0); a1 = fmaf(c1.y, v1[5], a1);
      a0 = fmaf(c1.z, v0[6], a0); a1 = fmaf(c1.z, v1[6], a1);
      a0 = fmaf(c1.w, v0[7], a0); a1 = fmaf(c1.w, v1[7], a1);
      a0 = fmaf(c2, v0[8], a0);   a1 = fmaf(c2, v1[8], a1);
      *reinterpret_cast<uint32_t*>(dst + coff[c]) = pack_bf16x2(a0, a1);
    }
    if (q != nullptr && by == 0) {   // chunk 0 of the forward matrix also emits q[i, o]
      const int i = i0 + r, o = o0 + 2 * tx;
      float q0 = 0.f, q1 = 0.f;
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        q0 = fmaf(sw[tp][r][2 * tx], sw[tp][r][2 * tx], q0);
        q1 = fmaf(sw[tp][r][2 * tx + 1], sw[tp][r][2 * tx + 1], q1);
      }
      if (i < p.I && o < p.O) q[static_cast<size_t>(i) * p.O + o] = q0;
      if (i < p.I && o + 1 < p.O) q[static_cast<size_t>(i) * p.O + o + 1] = q1;
    }
  }
}

__global__ void __launch_bounds__(256)
wprep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ adj,
             float* __restrict__ q, const WPrepParams p, const int chunk, const int fwd_chunks) {
  wprep_tile(w, fwd, adj, q, p, chunk, fwd_chunks, blockIdx.x, blockIdx.y);
}

// Every weight of a training iteration in ONE launch: the per-layer launches are latency-bound (~25 us each for a few
// hundred KB, 34 of them per iteration = 0.85 ms of a 19 ms step, profiles/r02n_graph_timeline_config2.txt).  A job is the
// argument set of one tbg_wprep call; CTA b works on job j with block_begin[j] <= b < block_begin[j+1].
struct WPrepJob {
  const float* w;
  __nv_bfloat16* fwd;
  __nv_bfloat16* adj;
  float* q;
  WPrepParams p;
  int chunk, fwd_chunks, tiles, blocks, block_begin, reserved;
};

__global__ void __launch_bounds__(256)
wprep_group_kernel(const WPrepJob* __restrict__ jobs, const int n_jobs) {
  __shared__ int sj;
  if (threadIdx.x == 0) {
    int lo = 0, hi = n_jobs - 1;                       // last job whose block_begin <= blockIdx.x
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].block_begin <= static_cast<int>(blockIdx.x)) lo = mid; else hi = mid - 1;
    }
    sj = lo;
  }
  __syncthreads();
  const WPrepJob& jb = jobs[sj];
  const int local = static_cast<int>(blockIdx.x) - jb.block_begin;
  wprep_tile(jb.w, jb.fwd, jb.adj, jb.q, jb.p, jb.chunk, jb.fwd_chunks, local % jb.tiles, local / jb.tiles);
}

// gw[tap,i,o] += coef * sum_c cf[c][tap] * gfwd_c[o,i]  (+ 2*coef^2*w*gq[i,o]).  One CTA per 32(i) x 8(o)
// weight tile: each thread owns one element, issues the loads of all (phase, tap) combinations
// back-to-back (lanes over the contiguous i axis of the gradient matrix, no barrier in the loop) and
// accumulates the nine master taps in registers; the result is transposed through shared memory so the
// read-modify-write of the o-contiguous master gradient touches whole 32-byte sectors.
__global__ void __launch_bounds__(256)
wfold_kernel(const float* __restrict__ gfwd, const float* __restrict__ gq, const float* __restrict__ w,
             float* __restrict__ gw, const WPrepParams p, const float* __restrict__ sv, const float* __restrict__ tv,
             const int nb, const int accumulate) {
  __shared__ float st[9][8][36];  // [tap][o][i]
  __shared__ __align__(16) float cff[36][12];
  const int tiles_o = p.Opad / 8;
  const int i0 = (blockIdx.x / tiles_o) * 32, o0 = (blockIdx.x % tiles_o) * 8;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int TT = p.fy.T * p.fx.T;
  const size_t Kf = static_cast<size_t>(TT) * p.Ipad;
  const int nph = p.fy.P * p.fx.P;
  const int taps = p.KH * p.KW;
  build_cf(p.fy, p.fx, p.KH, p.KW, p.coef, cff);
  __syncthreads();
  float acc[9];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) acc[tp] = 0.f;
  for (int pq = 0; pq < nph; ++pq) {
    const float* src = gfwd + (static_cast<size_t>(pq) * p.Opad + o0 + ty) * Kf + i0 + tx;
    int tu = 0;
    for (; tu + 4 <= TT; tu += 4) {     // four independent loads in flight per thread before the FMAs
      float g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) g[u] = __ldg(src + static_cast<size_t>(tu + u) * p.Ipad);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* cfr = cff[pq * TT + tu + u];
        const float4 c0 = *reinterpret_cast<const float4*>(cfr);
        const float4 c1 = *reinterpret_cast<const float4*>(cfr + 4);
        const float c2 = cfr[8];
        acc[0] = fmaf(c0.x, g[u], acc[0]); acc[1] = fmaf(c0.y, g[u], acc[1]); acc[2] = fmaf(c0.z, g[u], acc[2]);
        acc[3] = fmaf(c0.w, g[u], acc[3]); acc[4] = fmaf(c1.x, g[u], acc[4]); acc[5] = fmaf(c1.y, g[u], acc[5]);
        acc[6] = fmaf(c1.z, g[u], acc[6]); acc[7] = fmaf(c1.w, g[u], acc[7]); acc[8] = fmaf(c2, g[u], acc[8]);
      }
    }
    for (; tu < TT; ++tu) {
      const float g = __ldg(src + static_cast<size_t>(tu) * p.Ipad);
      const float* cfr = cff[pq * TT + tu];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) acc[tp] = fmaf(cfr[tp], g, acc[tp]);
    }
  }
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) st[tp][ty][tx] = acc[tp];
  __syncthreads();
  const int il = threadIdx.x >> 3, ol = threadIdx.x & 7;
  const int i = i0 + il, o = o0 + ol;
  if (i < p.I && o < p.O) {
    // dL/dq[i,o]: given, or formed here from the demodulation gradient t: sum_b s[b,i]^2 t[b,o]
    float gqv = 0.f;
    if (gq != nullptr) {
      gqv = __ldg(gq + static_cast<size_t>(i) * p.O + o);
    } else if (sv != nullptr) {
      for (int b = 0; b < nb; ++b) {
        const float s1 = __ldg(sv + static_cast<size_t>(b) * p.I + i);
        gqv = fmaf(s1 * s1, __ldg(tv + static_cast<size_t>(b) * p.O + o), gqv);
      }
    }
    gqv *= 2.f * p.coef * p.coef;
    const bool has_q = (gq != nullptr) || (sv != nullptr);
    for (int tp = 0; tp < taps; ++tp) {
      const size_t idx = (static_cast<size_t>(tp) * p.I + i) * p.O + o;
      float v = st[tp][ol][il];
      if (has_q) v = fmaf(gqv, __ldg(w + idx), v);
      gw[idx] = accumulate ? gw[idx] + v : v;
    }
  }
}

// Fold of a gradient held in the ADJOINT matrix layout of a plain (identity-table) geometry:
//   gw[tap,i,o] += coef * gadj[i, tap'*Opad + o]  (+ 2 coef^2 w[tap,i,o] * dL/dq[i,o]),  tap' = tap or, with
//   flip, the spatially mirrored tap (conv2d_transpose of upsample_conv_2d uses w[::-1, ::-1])
// (the role-swapped weight gradient of a transposed convolution lands in this layout).  o-contiguous on
// both sides: one thread per (i, o), loop over taps.
__global__ void __launch_bounds__(256)
wfold_adj_kernel(const float* __restrict__ gadj, const float* __restrict__ w, float* __restrict__ gw, int taps, int I,
                 int O, int Opad, float coef, const float* __restrict__ sv, const float* __restrict__ tv, int nb,
                 int flip, int accumulate) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= static_cast<long long>(I) * O) return;
  const int i = static_cast<int>(e / O), o = static_cast<int>(e % O);
  float gqv = 0.f;
  if (sv != nullptr) {
    for (int b = 0; b < nb; ++b) {
      const float s1 = __ldg(sv + static_cast<size_t>(b) * I + i);
      gqv = fmaf(s1 * s1, __ldg(tv + static_cast<size_t>(b) * O + o), gqv);
    }
    gqv *= 2.f * coef * coef;
  }
  for (int tp = 0; tp < taps; ++tp) {
    const size_t idx = (static_cast<size_t>(tp) * I + i) * O + o;
    const int ta = flip ? taps - 1 - tp : tp;   // the matrix holds the spatially flipped kernel
    float v = coef * __ldg(gadj + (static_cast<size_t>(i) * taps + ta) * Opad + o);
    if (sv != nullptr) v = fmaf(gqv, __ldg(w + idx), v);
    gw[idx] = accumulate ? gw[idx] + v : v;
  }
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_wfold_adj(const float* gadj, const float* w, float coef, int KH, int KW, int I, int O, int Opad,
                             float* gw, const float* s, const float* t, int nb, int flip, int accumulate, void* stream_v) {
  TBG_CHECK_ARG(gadj && gw, "tbg_wfold_adj: null pointer");
  TBG_CHECK_ARG(KH >= 1 && KH <= 3 && KW >= 1 && KW <= 3 && I >= 1 && O >= 1 && Opad >= O, "tbg_wfold_adj: bad shape");
  TBG_CHECK_ARG((s == nullptr) == (t == nullptr) && (!s || (w && nb >= 1)), "tbg_wfold_adj: (s, t, nb) need the master weight");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long n = static_cast<long long>(I) * O;
  wfold_adj_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(gadj, w, gw, KH * KW, I, O, Opad, coef, s, t, nb, flip, accumulate);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

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

// Validates one tbg_wprep argument set and fills the launch description shared by the single and the grouped entry.
static int make_wprep_job(WPrepJob& jb, const char* who, const float* w, const float* tables, float coef, int KH, int KW,
                          int I, int O, int Ipad, int Opad, void* fwd, void* adj, float* q, bool grouped) {
  TBG_CHECK_ARG(w && tables && fwd, "%s: null pointer", who);
  TBG_CHECK_ARG(KH >= 1 && KH <= 3 && KW >= 1 && KW <= 3, "%s: master kernel must be at most 3x3", who);
  TBG_CHECK_ARG(Ipad % 32 == 0 && Opad % 32 == 0 && Ipad >= I && Opad >= O && I >= 1 && O >= 1,
                "%s: bad channel counts I=%d O=%d Ipad=%d Opad=%d", who, I, O, Ipad, Opad);
  WPrepParams& p = jb.p;
  TBG_CHECK_ARG(load_tables(tables, p) == 0, "%s: malformed tables", who);
  TBG_CHECK_ARG(p.fy.K == KH && p.fx.K == KW, "%s: tables do not match the kernel size", who);
  p.KH = KH; p.KW = KW; p.I = I; p.O = O; p.Ipad = Ipad; p.Opad = Opad; p.coef = coef;
  const int ncf = p.fy.P * p.fx.P * p.fy.T * p.fx.T;
  const int nca = (p.ay.P > 0) ? p.ay.P * p.ax.P * p.ay.T * p.ax.T : 0;
  TBG_CHECK_ARG(ncf <= 36 && nca <= 36, "%s: too many (phase, tap) combinations", who);
  TBG_CHECK_ARG(nca == 0 || adj, "%s: adjoint tables without an adjoint output", who);
  // split the (phase, tap) combinations into equal chunks of at most 9 (one CTA each): every CTA then does the
  // same amount of work, so partial waves cost little, and even a 128 x 128 weight fills the SMs (a grouped launch is
  // filled by the other jobs: it keeps the larger chunks)
  const int tiles = (Ipad / 32) * (Opad / 32);
  const int nmax = ncf > nca ? ncf : nca;
  int chunk = nmax < 9 ? nmax : 9;
  if (!grouped && tiles * ((ncf + chunk - 1) / chunk + (nca + chunk - 1) / chunk) < 296 && chunk > 3) chunk = 3;
  jb.w = w;
  jb.fwd = reinterpret_cast<__nv_bfloat16*>(fwd);
  jb.adj = reinterpret_cast<__nv_bfloat16*>(adj);
  jb.q = q;
  jb.chunk = chunk;
  jb.fwd_chunks = (ncf + chunk - 1) / chunk;
  jb.tiles = tiles;
  jb.blocks = tiles * (jb.fwd_chunks + (nca + chunk - 1) / chunk);
  jb.block_begin = 0;
  jb.reserved = 0;
  return TBG_OK;
}

extern "C" int tbg_wprep(const float* w, const float* tables, float coef, int KH, int KW, int I, int O, int Ipad,
                         int Opad, void* fwd, void* adj, float* q, void* stream_v) {
  WPrepJob jb;
  const int rc = make_wprep_job(jb, "tbg_wprep", w, tables, coef, KH, KW, I, O, Ipad, Opad, fwd, adj, q, false);
  if (rc != TBG_OK) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  wprep_kernel<<<dim3(jb.tiles, jb.blocks / jb.tiles), 256, 0, stream>>>(jb.w, jb.fwd, jb.adj, jb.q, jb.p, jb.chunk,
                                                                          jb.fwd_chunks);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_wprep_job_bytes(void) { return static_cast<int>(sizeof(WPrepJob)); }

extern "C" int tbg_wprep_make_job(void* job_host, int block_begin, const float* w, const float* tables, float coef, int KH,
                                  int KW, int I, int O, int Ipad, int Opad, void* fwd, void* adj, float* q) {
  TBG_CHECK_ARG(job_host != nullptr && block_begin >= 0, "tbg_wprep_make_job: bad job buffer / block offset");
  WPrepJob jb;
  const int rc = make_wprep_job(jb, "tbg_wprep_make_job", w, tables, coef, KH, KW, I, O, Ipad, Opad, fwd, adj, q, true);
  if (rc != TBG_OK) return rc;
  jb.block_begin = block_begin;
  memcpy(job_host, &jb, sizeof(jb));
  return jb.blocks;                                   // > 0: CTAs of this job
}

extern "C" int tbg_wprep_group(const void* jobs_dev, int n_jobs, int total_blocks, void* stream_v) {
  TBG_CHECK_ARG(jobs_dev != nullptr && n_jobs >= 1 && total_blocks >= n_jobs, "tbg_wprep_group: bad job table");
  TBG_CHECK_ARG((reinterpret_cast<uintptr_t>(jobs_dev) & 15) == 0, "tbg_wprep_group: job table must be 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  wprep_group_kernel<<<total_blocks, 256, 0, stream>>>(reinterpret_cast<const WPrepJob*>(jobs_dev), n_jobs);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_wfold(const float* gfwd, const float* gq, const float* w, const float* tables, float coef, int KH,
                         int KW, int I, int O, int Ipad, int Opad, float* gw, const float* s, const float* t, int nb,
                         int accumulate, void* stream_v) {
  TBG_CHECK_ARG(gfwd && tables && gw, "tbg_wfold: null pointer");
  TBG_CHECK_ARG(!(gq || s) || w, "tbg_wfold: gq / (s, t) need the master weight");
  TBG_CHECK_ARG((s == nullptr) == (t == nullptr) && !(s && gq) && (!s || nb >= 1), "tbg_wfold: pass gq or (s, t, nb)");
  TBG_CHECK_ARG(KH >= 1 && KH <= 3 && KW >= 1 && KW <= 3, "tbg_wfold: master kernel must be at most 3x3");
  TBG_CHECK_ARG(Ipad % 32 == 0 && Opad % 32 == 0 && Ipad >= I && Opad >= O, "tbg_wfold: bad channel counts");
  WPrepParams p;
  TBG_CHECK_ARG(load_tables(tables, p) == 0, "tbg_wfold: malformed tables");
  TBG_CHECK_ARG(p.fy.P * p.fx.P * p.fy.T * p.fx.T <= 36, "tbg_wfold: too many (phase, tap) combinations");
  p.KH = KH; p.KW = KW; p.I = I; p.O = O; p.Ipad = Ipad; p.Opad = Opad; p.coef = coef;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  wfold_kernel<<<(Ipad / 32) * (Opad / 8), 256, 0, stream>>>(gfwd, gq, w, gw, p, s, t, nb, accumulate);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
