// Greedy Bahdanau-attention LSTM decoder of the frozen OCR head (forward + input-gradient BPTT),
// one CTA per sample, all decode steps in one launch.
//
// Per step (oracle/aster.py::attention_decoder; reached through AsterInferer.call,
// aster_ocr_utils/aster_inferer.py:28-37):
//   q = h Wq ; e_i = v . tanh(keys_i + q) ; a = softmax(e) ; ctx = sum_i a_i mem_i
//   gates = [emb[prev], ctx] W_ih + b + h W_hh ; (h, c) = LSTM(gates, c)
//   logits = [h, ctx] Wd + bd ; prev = argmax(logits)            (argmax feedback: no gradient)
// The head is frozen: backward returns only d/d mem and d/d keys (keys = mem Wm is a plain GEMM done
// by the caller).  Weights are bf16, streamed from L2 each step (2.4 MB per step per sample), fp32
// accumulation; 512 threads; latency-bound by construction (sequential in the step index).
#include "common.cuh"
#include "host_util.h"

namespace tbg {

static constexpr int DH = 256;    // LSTM hidden, attention units, embedding dim
static constexpr int DM = 512;    // memory (encoder) feature dim
static constexpr int DIN = DH + DM;  // LSTM input = [emb, ctx]; dense input = [h, ctx]
static constexpr int NC = 96;     // classes
static constexpr int TMAX = 64;   // max encoder length
static constexpr int NT = 512;    // threads

struct DecWeights {
  const __nv_bfloat16* wq;     // [DH][DH]        q[j]    = sum_k h[k]   wq[k][j]
  const __nv_bfloat16* wqT;    // [DH][DH]        dh[k]   = sum_j dq[j]  wqT[j][k]
  const float* v;              // [DH]
  const __nv_bfloat16* emb;    // [NC][DH]
  const __nv_bfloat16* wg;     // [DIN + DH][4*DH]  rows = [emb | ctx | h]   (W_ih stacked on W_hh; gates i,f,g,o)
  const __nv_bfloat16* wgT;    // [4*DH][DIN + DH]  its transpose
  const float* b;              // [4*DH]
  const __nv_bfloat16* wd;     // [DIN][NC]         rows = [h | ctx]
  const __nv_bfloat16* wdT;    // [NC][DIN]
  const float* bd;             // [NC]
};

__device__ __forceinline__ float bf(const __nv_bfloat16* p) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + __expf(-x)); }

// out[n] = sum_k x[k] * W[k][n] for a row-major bf16 matrix, all NT threads: thread (g, cv) owns 8
// consecutive columns (one 16-byte load per row) and a slice of K; partial sums meet in shared memory.
// Many independent 16-byte loads per thread keep the L2 stream busy (the decoder is latency-bound).
__device__ __forceinline__ void block_gemv(const float* __restrict__ x_s, int K, const __nv_bfloat16* __restrict__ W,
                                           int N, float* __restrict__ out_s, float* __restrict__ part_s) {
  const int ncv = N >> 3;
  const int groups = NT / ncv;
  const int cv = threadIdx.x % ncv, g = threadIdx.x / ncv;
  if (g < groups) {
    const int kper = (K + groups - 1) / groups;
    const int k0 = g * kper;
    const int k1 = min(K, k0 + kper);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const uint4* wp = reinterpret_cast<const uint4*>(W) + cv;
#pragma unroll 16
    for (int k = k0; k < k1; ++k) {
      const uint4 wv = __ldg(wp + static_cast<size_t>(k) * ncv);
      const float xk = x_s[k];
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(h2[j]);
        acc[2 * j] = fmaf(xk, f.x, acc[2 * j]);
        acc[2 * j + 1] = fmaf(xk, f.y, acc[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part_s[g * N + cv * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += NT) {
    float v = 0.f;
    for (int gg = 0; gg < groups; ++gg) v += part_s[gg * N + n];
    out_s[n] = v;
  }
  __syncthreads();
}

// saved per (b, step): a[T], ctx[DM], gates[4*DH] (post-activation), c[DH], h[DH], prev (int)
__global__ void __launch_bounds__(NT)
attn_decoder_fwd_kernel(const float* __restrict__ mem, const float* __restrict__ keys, const DecWeights W,
                        float* __restrict__ logits, float* __restrict__ sv_a, float* __restrict__ sv_ctx,
                        float* __restrict__ sv_gates, float* __restrict__ sv_c, float* __restrict__ sv_h,
                        int* __restrict__ sv_prev, int T, int steps) {
  // xin = [emb | ctx | h] feeds the stacked gate matrix; hc = [h | ctx] feeds the output dense
  __shared__ float xin_s[DIN + DH], hc_s[DIN], c_s[DH], q_s[DH], a_s[TMAX], gates_s[4 * DH], lg_s[NC];
  __shared__ float part_s[8 * NT];
  __shared__ int prev_s;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* memb = mem + static_cast<size_t>(b) * T * DM;
  const float* keyb = keys + static_cast<size_t>(b) * T * DH;
  float* emb_s = xin_s;
  float* ctx_s = xin_s + DH;
  float* h_s = xin_s + DIN;
  if (tid < DH) {
    h_s[tid] = 0.f;
    c_s[tid] = 0.f;
  }
  if (tid == 0) prev_s = 0;  // GO symbol
  __syncthreads();
  for (int st = 0; st < steps; ++st) {
    const size_t so = static_cast<size_t>(b) * steps + st;
    if (tid < DH) emb_s[tid] = bf(W.emb + prev_s * DH + tid);
    block_gemv(h_s, DH, W.wq, DH, q_s, part_s);                       // q = h Wq
    // ---- e_i = v . tanh(keys_i + q): 16 warps, each handles i = warp, warp+16, ... ----
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < T; i += NT / 32) {
        float acc = 0.f;
        for (int j = lane; j < DH; j += 32) acc = fmaf(__ldg(W.v + j), tanhf(__ldg(keyb + i * DH + j) + q_s[j]), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) a_s[i] = acc;
      }
    }
    __syncthreads();
    // ---- softmax over T (T <= 64: every thread reduces redundantly from shared memory) ----
    {
      float mx = -1e30f;
      for (int i = 0; i < T; ++i) mx = fmaxf(mx, a_s[i]);
      float sum = 0.f;
      for (int i = 0; i < T; ++i) sum += __expf(a_s[i] - mx);
      __syncthreads();
      if (tid < T) {
        const float a = __expf(a_s[tid] - mx) / sum;
        a_s[tid] = a;
        sv_a[so * T + tid] = a;
      }
    }
    __syncthreads();
    // ---- ctx = sum_i a_i mem_i ----
    {
      float acc = 0.f;
      for (int i = 0; i < T; ++i) acc = fmaf(a_s[i], __ldg(memb + i * DM + tid), acc);
      ctx_s[tid] = acc;
      sv_ctx[so * DM + tid] = acc;
    }
    __syncthreads();
    block_gemv(xin_s, DIN + DH, W.wg, 4 * DH, gates_s, part_s);      // [emb, ctx, h] [W_ih; W_hh]
    if (tid < DH) {
      const float ig = sigm(gates_s[tid] + __ldg(W.b + tid)), fg = sigm(gates_s[DH + tid] + __ldg(W.b + DH + tid)),
                  gg = tanhf(gates_s[2 * DH + tid] + __ldg(W.b + 2 * DH + tid)),
                  og = sigm(gates_s[3 * DH + tid] + __ldg(W.b + 3 * DH + tid));
      const float c = fg * c_s[tid] + ig * gg;
      const float h = og * tanhf(c);
      c_s[tid] = c;
      h_s[tid] = h;
      hc_s[tid] = h;
      float* sg = sv_gates + so * 4 * DH;
      sg[tid] = ig;
      sg[DH + tid] = fg;
      sg[2 * DH + tid] = gg;
      sg[3 * DH + tid] = og;
      sv_c[so * DH + tid] = c;
      sv_h[so * DH + tid] = h;
    }
    hc_s[DH + tid] = ctx_s[tid];
    if (tid == 0) sv_prev[so] = prev_s;
    __syncthreads();
    block_gemv(hc_s, DIN, W.wd, NC, lg_s, part_s);                   // logits = [h, ctx] Wd + bd
    if (tid < NC) {
      const float v = lg_s[tid] + __ldg(W.bd + tid);
      lg_s[tid] = v;
      logits[so * NC + tid] = v;
    }
    __syncthreads();
    if (tid == 0) {  // greedy feedback: first maximal class (torch.argmax semantics)
      int best = 0;
      float bv = lg_s[0];
      for (int i = 1; i < NC; ++i)
        if (lg_s[i] > bv) {
          bv = lg_s[i];
          best = i;
        }
      prev_s = best;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(NT)
attn_decoder_bwd_kernel(const float* __restrict__ mem, const float* __restrict__ keys, const DecWeights W,
                        const float* __restrict__ g_logits, const float* __restrict__ sv_a,
                        const float* __restrict__ sv_ctx, const float* __restrict__ sv_gates,
                        const float* __restrict__ sv_c, const float* __restrict__ sv_h,
                        float* __restrict__ g_mem, float* __restrict__ g_keys, int T, int steps) {
  __shared__ float dh_s[DH], dc_s[DH], dctx_s[DM], dg_s[4 * DH], gl_s[NC], da_s[TMAX], de_s[TMAX], q_s[DH], dq_s[DH],
      hprev_s[DH], dcat_s[DIN], dx_s[DIN + DH], tmp_s[DH];
  __shared__ float part_s[8 * NT];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* memb = mem + static_cast<size_t>(b) * T * DM;
  const float* keyb = keys + static_cast<size_t>(b) * T * DH;
  float* gmemb = g_mem + static_cast<size_t>(b) * T * DM;
  float* gkeyb = g_keys + static_cast<size_t>(b) * T * DH;
  for (int i = tid; i < T * DM; i += NT) gmemb[i] = 0.f;
  for (int i = tid; i < T * DH; i += NT) gkeyb[i] = 0.f;
  if (tid < DH) {
    dh_s[tid] = 0.f;
    dc_s[tid] = 0.f;
  }
  __syncthreads();
  for (int st = steps - 1; st >= 0; --st) {
    const size_t so = static_cast<size_t>(b) * steps + st;
    if (tid < NC) gl_s[tid] = __ldg(g_logits + so * NC + tid);
    if (tid >= 256 && tid < 256 + DH) {
      const int k = tid - 256;
      hprev_s[k] = (st > 0) ? __ldg(sv_h + (so - 1) * DH + k) : 0.f;
    }
    __syncthreads();
    block_gemv(gl_s, NC, W.wdT, DIN, dcat_s, part_s);               // d[h, ctx] from the logits
    // ---- LSTM cell backward ----
    if (tid < DH) {
      const float* sg = sv_gates + so * 4 * DH;
      const float ig = __ldg(sg + tid), fg = __ldg(sg + DH + tid), gg = __ldg(sg + 2 * DH + tid), og = __ldg(sg + 3 * DH + tid);
      const float cc = __ldg(sv_c + so * DH + tid);
      const float cp = (st > 0) ? __ldg(sv_c + (so - 1) * DH + tid) : 0.f;
      const float dh = dh_s[tid] + dcat_s[tid];
      const float tc = tanhf(cc);
      const float dog = dh * tc * og * (1.f - og);
      const float dc = dh * og * (1.f - tc * tc) + dc_s[tid];
      dg_s[tid] = dc * gg * ig * (1.f - ig);
      dg_s[DH + tid] = dc * cp * fg * (1.f - fg);
      dg_s[2 * DH + tid] = dc * ig * (1.f - gg * gg);
      dg_s[3 * DH + tid] = dog;
      dc_s[tid] = dc * fg;
    }
    __syncthreads();
    block_gemv(dg_s, 4 * DH, W.wgT, DIN + DH, dx_s, part_s);        // d[emb, ctx, h_prev] = dgates [W_ih; W_hh]^T
    block_gemv(hprev_s, DH, W.wq, DH, q_s, part_s);                 // recompute q = h_prev Wq
    dctx_s[tid] = dcat_s[DH + tid] + dx_s[DH + tid];
    __syncthreads();
    // ---- ctx = sum_i a_i mem_i : da_i = dctx . mem_i ; g_mem_i += a_i dctx ----
    {
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < T; i += NT / 32) {
        float acc = 0.f;
        for (int k = lane; k < DM; k += 32) acc = fmaf(dctx_s[k], __ldg(memb + i * DM + k), acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) da_s[i] = acc;
      }
      for (int i = 0; i < T; ++i) gmemb[i * DM + tid] += __ldg(sv_a + so * T + i) * dctx_s[tid];
    }
    __syncthreads();
    // ---- softmax backward: de_i = a_i (da_i - sum_j a_j da_j) ----
    {
      float dot = 0.f;
      for (int i = 0; i < T; ++i) dot = fmaf(__ldg(sv_a + so * T + i), da_s[i], dot);
      if (tid < T) de_s[tid] = __ldg(sv_a + so * T + tid) * (da_s[tid] - dot);
    }
    __syncthreads();
    // ---- e_i = v . tanh(keys_i + q): d pre_i[j] = de_i v_j (1 - tanh^2) ; g_keys_i += ; dq = sum_i ----
    if (tid < DH) {
      const float vj = __ldg(W.v + tid), qj = q_s[tid];
      float dq = 0.f;
      for (int i = 0; i < T; ++i) {
        const float th = tanhf(__ldg(keyb + i * DH + tid) + qj);
        const float dp = de_s[i] * vj * (1.f - th * th);
        gkeyb[i * DH + tid] += dp;
        dq += dp;
      }
      dq_s[tid] = dq;
    }
    __syncthreads();
    block_gemv(dq_s, DH, W.wqT, DH, tmp_s, part_s);                 // dq Wq^T
    if (tid < DH) dh_s[tid] = dx_s[DIN + tid] + tmp_s[tid];         // carried to step st-1
    __syncthreads();
  }
}

}  // namespace tbg

using namespace tbg;

static DecWeights to_dw(const tbg_dec_weights* w) {
  DecWeights d;
  d.wq = reinterpret_cast<const __nv_bfloat16*>(w->wq);
  d.wqT = reinterpret_cast<const __nv_bfloat16*>(w->wqT);
  d.v = reinterpret_cast<const float*>(w->v);
  d.emb = reinterpret_cast<const __nv_bfloat16*>(w->emb);
  d.wg = reinterpret_cast<const __nv_bfloat16*>(w->wg);
  d.wgT = reinterpret_cast<const __nv_bfloat16*>(w->wgT);
  d.b = reinterpret_cast<const float*>(w->b);
  d.wd = reinterpret_cast<const __nv_bfloat16*>(w->wd);
  d.wdT = reinterpret_cast<const __nv_bfloat16*>(w->wdT);
  d.bd = reinterpret_cast<const float*>(w->bd);
  return d;
}

extern "C" int tbg_attn_decoder_fwd(const float* mem, const float* keys, const tbg_dec_weights* w, float* logits,
                                    float* sv_a, float* sv_ctx, float* sv_gates, float* sv_c, float* sv_h,
                                    int* sv_prev, int B, int T, int steps, void* stream_v) {
  TBG_CHECK_ARG(mem && keys && w && logits && sv_a && sv_ctx && sv_gates && sv_c && sv_h && sv_prev,
                "tbg_attn_decoder_fwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && T >= 1 && T <= TMAX && steps >= 1, "tbg_attn_decoder_fwd: bad shape B=%d T=%d steps=%d", B, T, steps);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  attn_decoder_fwd_kernel<<<B, NT, 0, stream>>>(mem, keys, to_dw(w), logits, sv_a, sv_ctx, sv_gates, sv_c, sv_h,
                                                sv_prev, T, steps);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

extern "C" int tbg_attn_decoder_bwd(const float* mem, const float* keys, const tbg_dec_weights* w, const float* g_logits,
                                    const float* sv_a, const float* sv_ctx, const float* sv_gates, const float* sv_c,
                                    const float* sv_h, float* g_mem, float* g_keys, int B, int T, int steps,
                                    void* stream_v) {
  TBG_CHECK_ARG(mem && keys && w && g_logits && sv_a && sv_ctx && sv_gates && sv_c && sv_h && g_mem && g_keys,
                "tbg_attn_decoder_bwd: null pointer");
  TBG_CHECK_ARG(B >= 1 && T >= 1 && T <= TMAX && steps >= 1, "tbg_attn_decoder_bwd: bad shape B=%d T=%d steps=%d", B, T, steps);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  attn_decoder_bwd_kernel<<<B, NT, 0, stream>>>(mem, keys, to_dw(w), g_logits, sv_a, sv_ctx, sv_gates, sv_c, sv_h, g_mem,
                                                g_keys, T, steps);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
