// Implicit-GEMM convolution for sm_100a: TMA (tiled mode, shifted boxes, zero OOB fill) feeds
// 128-pixel x 64-channel activation tiles and N x 64 weight tiles into 128B-swizzled shared
// memory; one elected thread issues tcgen05.mma (M=128, N<=256, K=16) into double-buffered TMEM
// accumulators; four epilogue warps drain TMEM with tcgen05.ld and apply the StyleGAN2 epilogue
// (demod scale, noise, bias, leaky-ReLU, residual) before writing NHWC bf16/fp32.
//
// Reference behaviour covered: ModulatedConv2D.call (modulated_conv2d.py:66-122, non-fused
// algebra), upsample_conv_2d (upfirdn_2d_v2.py:65-103, as a 4-phase GEMM), Conv2D.call
// (conv.py:51-73), Noise.call (noise.py:12-22), BiasAct.call (bias_act.py:25-34), the residual
// merge of DiscriminatorBlock.call (discriminator.py:82).
#include <string.h>

#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct ConvKernelParams {
  int B;
  int bw, bh, bn;  // M-tile box (output pixels): bw*bh*bn <= 128 (any sizes; rows beyond the box are ignored)
  uint64_t tap_mask[4];  // per output phase: bit (ty*taps_w+tx) set = that tap's K blocks are computed
  int tiles_w, tiles_h, tiles_b, tiles_n;
  int block_n;  // columns per N tile (multiple of 32, <= 256)
  int n_total;
  int cin_chunks;  // Cin / 64
  int cin;
  int taps_h, taps_w;
  int in_off_h, in_off_w;  // = -pad
  int stride_h, stride_w;
  int up_h, up_w;
  int Ho, Wo;  // valid GEMM-row grid (tiles may overhang; rows outside are masked)
  int cout;
  int out_H, out_W;
  int stages;
  // epilogue
  const float* col_scale;
  const float* bias;
  const float* noise;
  const float* noise_strength;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* relu_mask;  // optional: output zeroed where this tensor (shape of out) is <= 0
  float res_scale;
  int res_first;  // 1: residual is added before the activation (ResNet unit), 0: after (D block)
  int act;
  float act_gain;
  int out_fp32;
  void* out;
};

static constexpr uint32_t kABytes = 128 * 128;  // 128 pixels x 64 bf16
static constexpr int kMaxStages = 8;

static constexpr int kEpiAffine = 1, kEpiRes = 2, kEpiMask = 4, kEpiF32 = 8, kEpiAll = 15;

template <int EPI>
__global__ void __launch_bounds__(384, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);

  const int stages = p.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(p.block_n) * 128u;
  uint8_t* smA = smem;
  uint8_t* smB = smem + stages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + stages * b_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kMaxStages;
  uint64_t* tfull = bars + 2 * kMaxStages;
  uint64_t* tempty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
  const int total_tiles = tiles_m * p.tiles_n;
  const int bw = p.bw, bh = p.bh, bn = p.bn;
  const uint32_t a_bytes = static_cast<uint32_t>(bw * bh * bn) * 128u;  // bytes one activation box delivers
  // output phase of an N tile (up-sampling geometries put the phases side by side along N)
  auto tile_mask = [&](int n_tile) -> uint64_t {
    const int ph = (p.up_h | p.up_w) ? (n_tile * p.block_n) / p.cout : 0;
    return p.tap_mask[ph & 3];
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.tiles_n;
        const int m_tile = tile / p.tiles_n;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tb = m_tile / (p.tiles_w * p.tiles_h);
        const int w_base = tw * bw * p.stride_w + p.in_off_w;
        const int h_base = th * bh * p.stride_h + p.in_off_h;
        const int n_base = tb * bn;
        const uint64_t mask = tile_mask(n_tile);
        for (int ty = 0; ty < p.taps_h; ++ty) {
          for (int tx = 0; tx < p.taps_w; ++tx) {
            if (!((mask >> (ty * p.taps_w + tx)) & 1ull)) continue;   // structurally zero weight block
            int kcol = (ty * p.taps_w + tx) * p.cin;
            for (int ch = 0; ch < p.cin_chunks; ++ch, kcol += 64) {
              mbar_wait(&empty[stage], phase ^ 1u);
              mbar_arrive_expect_tx(&full[stage], a_bytes + b_bytes);
              tma_load_4d(smA + stage * kABytes, &tmA, &full[stage], ch * 64, w_base + tx, h_base + ty, n_base);
              tma_load_2d(smB + stage * b_bytes, &tmB, &full[stage], kcol, n_tile * p.block_n);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.block_n), 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc_stage = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc_stage], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc_stage * p.block_n);
        const uint64_t mask = tile_mask(tile % p.tiles_n);
        const int k_blocks = __popcll(mask) * p.cin_chunks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + stage * kABytes);
          const uint32_t b_addr = smem_u32(smB + stage * b_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == k_blocks - 1) umma_commit(&tfull[acc_stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    // Eight warps: warp e may read TMEM lanes 32*(e%4).. (its quadrant = 32 tile rows); the two warps of a quadrant take
    // alternate 32-column groups.  One warp per scheduler runs this code latency-bound (ncu: ~400 dependent instructions
    // per 32 x 32 block with every feature tested at run time, profiles/r02q_igemm_epilogue.txt), so the feature set is a
    // template parameter (bit mask): kEpiAffine = column scale / noise / bias / activation, kEpiRes = residual,
    // kEpiMask = ReLU-backward mask, kEpiF32 = fp32 output; 0 = bare bf16 store.  Instantiated: the combinations the
    // training step uses (0, affine, affine + residual, mask, residual + mask) and the general one.
    const int e = warp - 4;
    const int quad = e & 3, grp = e >> 2;
    const int r = quad * 32 + lane;
    const int w_in = r % bw;
    const int h_in = (r / bw) % bh;
    const int n_in = r / (bw * bh);
    const float nstr = ((EPI & kEpiAffine) && p.noise != nullptr) ? __ldg(p.noise_strength) : 0.f;
    const bool has_up = (p.up_h | p.up_w) != 0;
    const int nj = p.block_n / 32;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int n_tile = tile % p.tiles_n;
      const int m_tile = tile / p.tiles_n;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tb = m_tile / (p.tiles_w * p.tiles_h);
      const int b = tb * bn + n_in;
      const int ho = th * bh + h_in;
      const int wo = tw * bw + w_in;
      const bool valid = (n_in < bn) && (b < p.B) && (ho < p.Ho) && (wo < p.Wo);
      const int acc_stage = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc_stage], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc_stage * p.block_n);
      for (int j = grp; j < nj; j += 2) {
        const int col0 = n_tile * p.block_n + j * 32;
        uint32_t v[32];
        tmem_ld_32x32(t_row + j * 32, v);
        tmem_ld_wait();
        if (!valid || col0 >= p.n_total) continue;
        int c0 = col0, oy = ho, ox = wo;
        if (has_up) {
          const int ph = col0 / p.cout;
          c0 = col0 - ph * p.cout;
          const int py = p.up_w ? (ph >> 1) : ph;
          const int px = p.up_w ? (ph & 1) : 0;
          oy = p.up_h ? 2 * ho + py : ho;
          ox = p.up_w ? 2 * wo + px : wo;
        }
        const size_t pix = (static_cast<size_t>(b) * p.out_H + oy) * p.out_W + ox;
        const size_t off = pix * p.cout + c0;
        if (EPI == 0) {
          uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + off);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            pk.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            pk.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            pk.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            pk.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            o[g] = pk;
          }
          continue;
        }
        const float nz = ((EPI & kEpiAffine) && p.noise != nullptr) ? __ldg(p.noise + pix) * nstr : 0.f;
        const float* cs = ((EPI & kEpiAffine) && p.col_scale) ? p.col_scale + static_cast<size_t>(b) * p.cout + c0 : nullptr;
        const float* bs = ((EPI & kEpiAffine) && p.bias) ? p.bias + c0 : nullptr;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]);
          if (cs) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(cs + g * 8));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(cs + g * 8 + 4));
            f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
            f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
          }
          if (EPI & kEpiAffine) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += nz;
          }
          if (bs) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs + g * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bs + g * 8 + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          float rres[8];
          if ((EPI & kEpiRes) && p.residual) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p.residual + off + g * 8));
            const __nv_bfloat162* rh = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 rf = __bfloat1622float2(rh[i]);
              rres[2 * i] = rf.x;
              rres[2 * i + 1] = rf.y;
            }
            if (p.res_first) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = (f[i] + rres[i]) * p.res_scale;
            }
          }
          if (EPI & kEpiAffine) {
            if (p.act == 1) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = (f[i] > 0.f ? f[i] : 0.2f * f[i]);
            } else if (p.act == 2) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] *= p.act_gain;
          }
          if ((EPI & kEpiRes) && p.residual && !p.res_first) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = (f[i] + rres[i]) * p.res_scale;
          }
          if ((EPI & kEpiMask) && p.relu_mask) {   // gradient of a ReLU whose output is relu_mask (fused ReLU backward)
            const uint4 mv = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + off + g * 8));
            const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 mf = __bfloat1622float2(mh[i]);
              if (!(mf.x > 0.f)) f[2 * i] = 0.f;
              if (!(mf.y > 0.f)) f[2 * i + 1] = 0.f;
            }
          }
          if ((EPI & kEpiF32) && p.out_fp32) {
            float* o = reinterpret_cast<float*>(p.out) + off + g * 8;
            *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off + g * 8;
            uint4 pk;
            pk.x = pack_bf16x2(f[0], f[1]);
            pk.y = pack_bf16x2(f[2], f[3]);
            pk.z = pack_bf16x2(f[4], f[5]);
            pk.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(o) = pk;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc_stage]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

Tuning g_tuning;

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_conv2d_igemm(const tbg_conv_args* a, void* stream_v) {
  if (!a) return set_error(TBG_ERR_INVALID_ARG, "tbg_conv2d_igemm: null args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TBG_CHECK_ARG(a->x && a->w && a->out, "tbg_conv2d_igemm: null tensor pointer");
  TBG_CHECK_ARG(a->B >= 1 && a->H >= 1 && a->W >= 1, "tbg_conv2d_igemm: bad input shape B=%d H=%d W=%d", a->B, a->H, a->W);
  TBG_CHECK_ARG(a->Cin >= 64 && a->Cin % 64 == 0, "tbg_conv2d_igemm: Cin=%d must be a multiple of 64", a->Cin);
  TBG_CHECK_ARG(a->Ho >= 1 && a->Wo >= 1, "tbg_conv2d_igemm: bad output grid Ho=%d Wo=%d", a->Ho, a->Wo);
  TBG_CHECK_ARG(a->cout >= 32 && a->cout % 32 == 0, "tbg_conv2d_igemm: cout=%d must be a multiple of 32", a->cout);
  TBG_CHECK_ARG((a->up_h == 0 || a->up_h == 1) && (a->up_w == 0 || a->up_w == 1), "tbg_conv2d_igemm: up_h/up_w must be 0 or 1");
  TBG_CHECK_ARG(a->n_total == a->cout * (1 + a->up_h) * (1 + a->up_w),
                "tbg_conv2d_igemm: n_total=%d inconsistent with cout=%d up=(%d,%d)", a->n_total, a->cout, a->up_h, a->up_w);
  TBG_CHECK_ARG(!(a->up_h && a->stride_h != 1) && !(a->up_w && a->stride_w != 1), "tbg_conv2d_igemm: up with stride on one axis");
  TBG_CHECK_ARG(a->act >= 0 && a->act <= 2, "tbg_conv2d_igemm: act must be 0 (linear), 1 (lrelu) or 2 (relu)");
  TBG_CHECK_ARG(a->taps_h >= 1 && a->taps_w >= 1 && a->taps_h <= 8 && a->taps_w <= 8 && a->taps_h * a->taps_w <= 36,
                "tbg_conv2d_igemm: bad taps");
  TBG_CHECK_ARG((a->stride_h == 1 || a->stride_h == 2) && (a->stride_w == 1 || a->stride_w == 2),
                "tbg_conv2d_igemm: strides must be 1 or 2");
  TBG_CHECK_ARG(!a->noise || a->noise_strength, "tbg_conv2d_igemm: noise without noise_strength");
  TBG_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
                "tbg_conv2d_igemm: tensors must be 16-byte aligned");
  TBG_CHECK_ARG((reinterpret_cast<uintptr_t>(a->col_scale) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->relu_mask) & 15) == 0,
                "tbg_conv2d_igemm: col_scale / bias / residual must be 16-byte aligned (vector loads in the epilogue)");

  // 3x3 stride-1 SAME convolutions on grids of 16 x 16 pixel blocks: one activation halo box per 64-channel block
  // shared by the nine taps (csrc/conv_halo.cu)
  if (g_tuning.conv_halo && conv_halo_applicable(a)) return conv_halo_launch(a, stream);

  ConvKernelParams p{};
  p.B = a->B;
  // M-tile box: bw x bh x bn output pixels (<= 128) maximising the fraction of useful GEMM rows;
  // ties go to the widest box (longest contiguous TMA rows).  Power-of-two grids get full tiles.
  int bw = 1, bh = 1, bn = 1;
  {
    double best = -1.0;
    const int wmax = a->Wo < 128 ? a->Wo : 128;
    for (int cw = wmax; cw >= 1; --cw) {
      if (cw * a->stride_w > 256) continue;
      const int tw = (a->Wo + cw - 1) / cw;
      const int hmax = (128 / cw) < a->Ho ? (128 / cw) : a->Ho;
      for (int ch = hmax; ch >= 1; --ch) {
        if (ch * a->stride_h > 256) continue;
        const int th = (a->Ho + ch - 1) / ch;
        int cn = 128 / (cw * ch);
        if (cn > a->B) cn = a->B;
        if (cn > 1 && (tw > 1 || th > 1)) cn = 1;      // batch several images per tile only when one tile covers an image
        if (cn < 1) cn = 1;
        const int tb = (a->B + cn - 1) / cn;
        const double eff = (double)a->Wo * a->Ho * a->B / ((double)tw * th * tb * 128.0);
        if (eff > best + 1e-9) {
          best = eff;
          bw = cw; bh = ch; bn = cn;
        }
      }
    }
  }
  p.bw = bw;
  p.bh = bh;
  p.bn = bn;
  p.tiles_w = (a->Wo + bw - 1) / bw;
  p.tiles_h = (a->Ho + bh - 1) / bh;
  p.Ho = a->Ho;
  p.Wo = a->Wo;
  p.tiles_b = (a->B + bn - 1) / bn;
  // N tile: largest of 256/128/64/32 that divides n_total (keeps every tile full)
  int block_n = 256;
  while (block_n > 32 && (a->n_total % block_n) != 0) block_n >>= 1;
  if (a->n_total % block_n != 0) block_n = 32;
  // prefer >= 2 N tiles' worth of parallelism only when M tiles are scarce
  const int tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
  while (block_n > 64 && tiles_m * (a->n_total / block_n) < num_sms() && (a->n_total % (block_n / 2)) == 0) block_n >>= 1;
  // up-sampling geometries: an N tile must not straddle two output phases (tap masks are per phase)
  if (a->up_h | a->up_w)
    while (block_n > 32 && (block_n > a->cout || a->cout % block_n != 0)) block_n >>= 1;
  p.block_n = block_n;
  {
    const uint64_t all = (1ull << (a->taps_h * a->taps_w)) - 1ull;   // taps_h * taps_w <= 36
    const int nph = (1 + a->up_h) * (1 + a->up_w);
    for (int i = 0; i < 4; ++i) {
      p.tap_mask[i] = (a->tap_mask[i] ? a->tap_mask[i] : all) & all;
      if (i < nph)
        TBG_CHECK_ARG(p.tap_mask[i] != 0, "tbg_conv2d_igemm: phase %d has no taps", i);
    }
  }
  p.tiles_n = (a->n_total + block_n - 1) / block_n;
  p.n_total = a->n_total;
  p.cin = a->Cin;
  p.cin_chunks = a->Cin / 64;
  p.taps_h = a->taps_h;
  p.taps_w = a->taps_w;
  p.in_off_h = -a->pad_h;
  p.in_off_w = -a->pad_w;
  p.stride_h = a->stride_h;
  p.stride_w = a->stride_w;
  p.up_h = a->up_h;
  p.up_w = a->up_w;
  p.cout = a->cout;
  p.out_H = a->up_h ? 2 * a->Ho : a->Ho;
  p.out_W = a->up_w ? 2 * a->Wo : a->Wo;
  p.col_scale = a->col_scale;
  p.bias = a->bias;
  p.noise = a->noise;
  p.noise_strength = a->noise_strength;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  p.relu_mask = reinterpret_cast<const __nv_bfloat16*>(a->relu_mask);
  p.res_scale = a->res_scale;
  p.res_first = a->res_first;
  p.act = a->act;
  p.act_gain = a->act_gain;
  p.out_fp32 = a->out_fp32;
  p.out = a->out;

  const uint32_t b_bytes = static_cast<uint32_t>(block_n) * 128u;
  const uint32_t stage_bytes = kABytes + b_bytes;
  const uint32_t budget = 227u * 1024u - 1024u /*align slack*/ - 256u /*barriers*/;
  int stages = static_cast<int>(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + 1024 + 256;

  // ---- tensor maps ----
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    // with element stride s the box spans b*s source elements and delivers b of them.  (A parity-split 5-D view of x
    // that turns every tap into a dense box measured the same, profiles/r02p_perf_shapes.log, and was dropped.)
    const uint32_t box[4] = {64, (uint32_t)(bw * a->stride_w), (uint32_t)(bh * a->stride_h), (uint32_t)bn};
    const uint32_t estr[4] = {1, (uint32_t)a->stride_w, (uint32_t)a->stride_h, 1};
    int rc = encode_tmap_bf16(&tmA, a->x, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t K = (uint64_t)a->taps_h * a->taps_w * a->Cin;
    const uint64_t dims[2] = {K, (uint64_t)a->n_total};
    const uint64_t strides[2] = {0, K * 2};
    const uint32_t box[2] = {64, (uint32_t)block_n};
    int rc = encode_tmap_bf16(&tmB, a->w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
#define TBG_IGEMM_ATTR(E) \
  TBG_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    TBG_IGEMM_ATTR(0);
    TBG_IGEMM_ATTR(kEpiAffine);
    TBG_IGEMM_ATTR(kEpiAffine | kEpiRes);
    TBG_IGEMM_ATTR(kEpiMask);
    TBG_IGEMM_ATTR(kEpiRes | kEpiMask);
    TBG_IGEMM_ATTR(kEpiAll);
#undef TBG_IGEMM_ATTR
    attr_set = true;
  }
  const int total_tiles = tiles_m * p.tiles_n;
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  // an activation with gain 1 and no scale / noise / bias still needs the affine path only when act != 0
  const bool affine = a->col_scale || a->noise || a->bias || a->act != 0 || a->act_gain != 1.f;
  const int need = (affine ? kEpiAffine : 0) | (a->residual ? kEpiRes : 0) | (a->relu_mask ? kEpiMask : 0) |
                   (a->out_fp32 ? kEpiF32 : 0);
#define TBG_IGEMM_LAUNCH(E) conv_igemm_kernel<E><<<grid, 384, smem_bytes, stream>>>(tmA, tmB, p)
  switch (need) {
    case 0: TBG_IGEMM_LAUNCH(0); break;
    case kEpiAffine: TBG_IGEMM_LAUNCH(kEpiAffine); break;
    case kEpiAffine | kEpiRes: TBG_IGEMM_LAUNCH(kEpiAffine | kEpiRes); break;
    case kEpiMask: TBG_IGEMM_LAUNCH(kEpiMask); break;
    case kEpiRes | kEpiMask: TBG_IGEMM_LAUNCH(kEpiRes | kEpiMask); break;
    default: TBG_IGEMM_LAUNCH(kEpiAll); break;
  }
#undef TBG_IGEMM_LAUNCH
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

// Explicit tuning switches (tests, perf scripts); the library reads no environment variables.
extern "C" int tbg_set_tuning(const char* key, int value) {
  TBG_CHECK_ARG(key != nullptr, "tbg_set_tuning: null key");
  if (!strcmp(key, "conv_halo")) g_tuning.conv_halo = value != 0;
  else if (!strcmp(key, "wgrad_staged")) g_tuning.wgrad_staged = value != 0;
  else if (!strcmp(key, "wgrad_items_per_sm")) g_tuning.wgrad_items_per_sm = value;
  else if (!strcmp(key, "lstm_cluster")) g_tuning.lstm_cluster = value != 0;
  else if (!strcmp(key, "halo_a_stages")) g_tuning.halo_a_stages = value < 2 ? 2 : (value > 3 ? 3 : value);
  else if (!strcmp(key, "halo_b_stages")) g_tuning.halo_b_stages = value < 2 ? 2 : (value > 8 ? 8 : value);
  else if (!strcmp(key, "halo_staged")) g_tuning.halo_staged = value != 0;
  else if (!strcmp(key, "wgrad_halo")) g_tuning.wgrad_halo = value != 0;
  else if (!strcmp(key, "halo_cta2")) g_tuning.halo_cta2 = value != 0;
  else return set_error(TBG_ERR_INVALID_ARG, "tbg_set_tuning: unknown key '%s'", key);
  return TBG_OK;
}

extern "C" int tbg_get_tuning(const char* key) {
  if (!key) return -1;
  if (!strcmp(key, "conv_halo")) return g_tuning.conv_halo;
  if (!strcmp(key, "wgrad_staged")) return g_tuning.wgrad_staged;
  if (!strcmp(key, "wgrad_items_per_sm")) return g_tuning.wgrad_items_per_sm;
  if (!strcmp(key, "lstm_cluster")) return g_tuning.lstm_cluster;
  if (!strcmp(key, "halo_a_stages")) return g_tuning.halo_a_stages;
  if (!strcmp(key, "halo_b_stages")) return g_tuning.halo_b_stages;
  if (!strcmp(key, "halo_staged")) return g_tuning.halo_staged;
  if (!strcmp(key, "wgrad_halo")) return g_tuning.wgrad_halo;
  if (!strcmp(key, "halo_cta2")) return g_tuning.halo_cta2;
  return -1;
}
