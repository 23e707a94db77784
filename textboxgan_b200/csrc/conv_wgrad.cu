// Weight-gradient of the implicit-GEMM convolution for sm_100a.
//
//   gw[n, tap, c] += sum_pixels gy[pixel, n] * x[pixel*stride - pad + tap, c]
//
// GEMM view: M = 128 output channels (from gy), N = a block of input channels, K = pixels.  Both
// operands arrive pixel-major ([pixels][64 channels] TMA boxes, 128B swizzle), i.e. "MN-major" for
// tcgen05.mma, so the instruction descriptor sets a_major = b_major = MN and the shared-memory
// descriptors use LBO = bytes between 64-channel boxes, SBO = 1024 (8 pixel rows).  One CTA keeps
// up to 512 TMEM columns of accumulators = several (tap, channel-block) sub-tiles that share each
// gy tile, walks a contiguous range of pixel blocks (split-K across CTAs) and reduces into fp32
// gw with vector atomics.
//
// Replaces cuDNN's backward-filter behind tape.gradient (training_step.py:224-235) for
// ModulatedConv2D / Conv2D weights (modulated_conv2d.py:63, conv.py:49).
#include <stdlib.h>

#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct WgradParams {
  int B;
  int bw_log2, bh_log2, bn_log2;  // pixel box, product = P
  int p_log2;                     // log2(P), P in {32, 64}
  int tiles_w, tiles_h, tiles_b;  // pixel tiles
  int m_tiles;                    // ceil(n_total / 128)
  int n_total, cout, up_h, up_w;
  int cin, block_c, c_tiles;      // N tiling of input channels
  int taps_h, taps_w;
  int subtiles;                   // taps * c_tiles
  int G;                          // sub-tiles (accumulators) per work item
  int groups;                     // ceil(subtiles / G)
  int splits;                     // split-K factor over pixel tiles
  int in_off_h, in_off_w, stride_h, stride_w;
  int stages;
  int ktot;                       // taps * cin
  float* gw;
  int staged;   // 1: vector atomics go out line-coalesced through shared memory (TBG_WGRAD_STAGED=0: one row per thread)
};

static constexpr int kWgMaxStages = 8;
// Epilogue staging (as in conv_igemm.cu): each epilogue warp transposes 32 accumulator rows x 64 fp32 columns
// through shared memory so that one red.global.add.v4 instruction covers two rows x 256 contiguous bytes
// instead of 16 bytes of 32 different rows.
static constexpr uint32_t kWgStgRow = 256 + 16;
static constexpr uint32_t kWgStgBytes = 4 * 32 * kWgStgRow;

__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmGY, const __grid_constant__ CUtensorMap tmX,
                  const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);

  const int P = 1 << p.p_log2;
  const uint32_t chunk_bytes = static_cast<uint32_t>(P) * 128u;       // one [P pixels][64 ch] box
  const uint32_t a_bytes = 2u * chunk_bytes;                            // 128 output channels
  const uint32_t sub_bytes = static_cast<uint32_t>(p.block_c / 64) * chunk_bytes;
  const uint32_t b_bytes = static_cast<uint32_t>(p.G) * sub_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int stages = p.stages;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgMaxStages;
  uint64_t* tfull = bars + 2 * kWgMaxStages;
  uint64_t* tempty = bars + 2 * kWgMaxStages + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 2);
  uint8_t* stg_base = reinterpret_cast<uint8_t*>(bars + 2 * kWgMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmGY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 128);
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

  const int bw = 1 << p.bw_log2, bh = 1 << p.bh_log2, bn = 1 << p.bn_log2;
  const int k_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
  const int total_items = p.m_tiles * p.groups * p.splits;

  // item -> (m_tile, group, split); split fastest so neighbouring CTAs share weights' target tile
  auto item_decode = [&](int item, int& m_tile, int& group, int& kt0, int& kt1) {
    const int split = item % p.splits;
    const int rest = item / p.splits;
    group = rest % p.groups;
    m_tile = rest / p.groups;
    const int per = (k_tiles + p.splits - 1) / p.splits;
    kt0 = split * per;
    kt1 = kt0 + per;
    if (kt1 > k_tiles) kt1 = k_tiles;
  };

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int m_tile, group, kt0, kt1;
        item_decode(item, m_tile, group, kt0, kt1);
        const int s_begin = group * p.G;
        int s_end = s_begin + p.G;
        if (s_end > p.subtiles) s_end = p.subtiles;
        for (int kt = kt0; kt < kt1; ++kt) {
          const int tw = kt % p.tiles_w;
          const int th = (kt / p.tiles_w) % p.tiles_h;
          const int tb = kt / (p.tiles_w * p.tiles_h);
          mbar_wait(&empty[stage], phase ^ 1u);
          const uint32_t tx_bytes = a_bytes + static_cast<uint32_t>(s_end - s_begin) * sub_bytes;
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          uint8_t* sa = smem + stage * stage_bytes;
          // gy: two 64-channel boxes
          for (int j = 0; j < 2; ++j) {
            const int n0 = m_tile * 128 + j * 64;
            int c0 = n0, wy = th * bh, wx = tw * bw;
            if (p.up_h | p.up_w) {
              const int ph = n0 / p.cout;
              c0 = n0 - ph * p.cout;
              const int py = p.up_w ? (ph >> 1) : ph;
              const int px = p.up_w ? (ph & 1) : 0;
              if (p.up_h) wy = 2 * wy + py;
              if (p.up_w) wx = 2 * wx + px;
              if (n0 >= p.n_total) c0 = p.cout;  // past the last phase: force OOB -> zero fill
            }
            tma_load_4d(sa + j * chunk_bytes, &tmGY, &full[stage], c0, wx, wy, tb * bn);
          }
          // x: one set of boxes per (tap, channel block) sub-tile
          uint8_t* sb = sa + a_bytes;
          for (int s = s_begin; s < s_end; ++s) {
            const int tap = s / p.c_tiles;
            const int ct = s - tap * p.c_tiles;
            const int ty = tap / p.taps_w;
            const int tx = tap - ty * p.taps_w;
            const int hx = th * bh * p.stride_h + p.in_off_h + ty;
            const int wx = tw * bw * p.stride_w + p.in_off_w + tx;
            for (int jc = 0; jc < p.block_c / 64; ++jc) {
              tma_load_4d(sb + (s - s_begin) * sub_bytes + jc * chunk_bytes, &tmX, &full[stage],
                          ct * p.block_c + jc * 64, wx, hx, tb * bn);
            }
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.block_c), 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        int m_tile, group, kt0, kt1;
        item_decode(item, m_tile, group, kt0, kt1);
        const int s_begin = group * p.G;
        int s_end = s_begin + p.G;
        if (s_end > p.subtiles) s_end = p.subtiles;
        mbar_wait(tempty, (it & 1) ^ 1u);
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t b_addr = a_addr + a_bytes;
          for (int s = 0; s < s_end - s_begin; ++s) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(s * p.block_c);
            for (int k = 0; k < P / 16; ++k) {
              const uint64_t da = umma_smem_desc_sw128(a_addr + k * 2048, chunk_bytes, 1024);
              const uint64_t db = umma_smem_desc_sw128(b_addr + s * sub_bytes + k * 2048, chunk_bytes, 1024);
              umma_bf16(d_tmem, da, db, idesc, (kt > kt0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull);
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int m_tile, group, kt0, kt1;
      item_decode(item, m_tile, group, kt0, kt1);
      const int s_begin = group * p.G;
      int s_end = s_begin + p.G;
      if (s_end > p.subtiles) s_end = p.subtiles;
      mbar_wait(tfull, it & 1);
      tc_fence_after();
      uint8_t* const stg = stg_base + e * (32 * kWgStgRow);
      for (int s = s_begin; s < s_end; ++s) {
        const int tap = s / p.c_tiles;
        const int ct = s - tap * p.c_tiles;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(e * 32) << 16) +
                               static_cast<uint32_t>((s - s_begin) * p.block_c);
        for (int j = 0; j < p.block_c / 32; ++j) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + j * 32, v);
          tmem_ld_wait();
          if (!p.staged) {      // round-1 path: every thread reduces 16-byte pieces of its own output-channel row
            const int nn = m_tile * 128 + e * 32 + lane;
            if (nn < p.n_total && kt1 > kt0) {
              float* dst = p.gw + static_cast<size_t>(nn) * p.ktot + tap * p.cin + ct * p.block_c + j * 32;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4),
                             "f"(__uint_as_float(v[g * 4])), "f"(__uint_as_float(v[g * 4 + 1])),
                             "f"(__uint_as_float(v[g * 4 + 2])), "f"(__uint_as_float(v[g * 4 + 3]))
                             : "memory");
            }
            continue;
          }
          float4* srow = reinterpret_cast<float4*>(stg + lane * kWgStgRow + (j & 1) * 128);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            srow[g] = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                  __uint_as_float(v[g * 4 + 3]));
          const bool last = (j + 1 == p.block_c / 32);
          if ((j & 1) || last) {
            // flush 64 (or the last 32) columns of the warp's 32 rows
            __syncwarp();
            const int cols = (j & 1) ? 64 : 32;
            const int lanes_per_row = cols / 4;               // 16 or 8 lanes x 16 bytes
            const int rows_per_pass = 32 / lanes_per_row;
            const int sub = lane % lanes_per_row;
            const int jc0 = (j & 1) ? j - 1 : j;              // first 32-column block of this chunk
            for (int r0 = 0; r0 < 32; r0 += rows_per_pass) {
              const int rr = r0 + lane / lanes_per_row;
              const int nn = m_tile * 128 + e * 32 + rr;
              if (nn < p.n_total && kt1 > kt0) {
                const float4 val = *reinterpret_cast<const float4*>(stg + rr * kWgStgRow + sub * 16);
                float* dst = p.gw + static_cast<size_t>(nn) * p.ktot + tap * p.cin + ct * p.block_c + jc0 * 32 + sub * 4;
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(val.x), "f"(val.y),
                             "f"(val.z), "f"(val.w)
                             : "memory");
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace tbg

using namespace tbg;

extern "C" int tbg_conv2d_wgrad(const tbg_wgrad_args* a, void* stream_v) {
  if (!a) return set_error(TBG_ERR_INVALID_ARG, "tbg_conv2d_wgrad: null args");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  TBG_CHECK_ARG(a->x && a->gy && a->gw, "tbg_conv2d_wgrad: null tensor pointer");
  TBG_CHECK_ARG(a->Cin >= 64 && a->Cin % 64 == 0, "tbg_conv2d_wgrad: Cin=%d must be a multiple of 64", a->Cin);
  TBG_CHECK_ARG(a->cout >= 32 && a->cout % 32 == 0, "tbg_conv2d_wgrad: cout=%d must be a multiple of 32", a->cout);
  TBG_CHECK_ARG(a->Ho >= 1 && a->Wo >= 1, "tbg_conv2d_wgrad: bad grid Ho=%d Wo=%d", a->Ho, a->Wo);
  TBG_CHECK_ARG((a->up_h == 0 || a->up_h == 1) && (a->up_w == 0 || a->up_w == 1), "tbg_conv2d_wgrad: up_h/up_w must be 0 or 1");
  TBG_CHECK_ARG(a->n_total == a->cout * (1 + a->up_h) * (1 + a->up_w), "tbg_conv2d_wgrad: n_total inconsistent");
  TBG_CHECK_ARG(!(a->up_h && a->stride_h != 1) && !(a->up_w && a->stride_w != 1), "tbg_conv2d_wgrad: up with stride on one axis");
  TBG_CHECK_ARG((a->stride_h == 1 || a->stride_h == 2) && (a->stride_w == 1 || a->stride_w == 2),
                "tbg_conv2d_wgrad: strides must be 1 or 2");
  TBG_CHECK_ARG((a->up_h | a->up_w) == 0 || a->cout % 64 == 0, "tbg_conv2d_wgrad: up needs cout %% 64 == 0");

  // 3x3 stride-1 SAME convolutions on large grids: one x halo box per 64-channel block shared by up to four taps
  if (g_tuning.wgrad_halo && wgrad_halo_applicable(a)) return wgrad_halo_launch(a, stream);

  WgradParams p{};
  p.B = a->B;
  p.n_total = a->n_total;
  p.cout = a->cout;
  p.up_h = a->up_h;
  p.up_w = a->up_w;
  p.cin = a->Cin;
  p.block_c = a->Cin < 256 ? a->Cin : 256;
  if (a->Cin % p.block_c != 0) p.block_c = 64;
  p.c_tiles = a->Cin / p.block_c;
  p.taps_h = a->taps_h;
  p.taps_w = a->taps_w;
  p.subtiles = a->taps_h * a->taps_w * p.c_tiles;
  const int gmax = 512 / p.block_c;
  p.groups = (p.subtiles + gmax - 1) / gmax;
  p.G = (p.subtiles + p.groups - 1) / p.groups;
  p.groups = (p.subtiles + p.G - 1) / p.G;
  p.m_tiles = (a->n_total + 127) / 128;
  p.in_off_h = -a->pad_h;
  p.in_off_w = -a->pad_w;
  p.stride_h = a->stride_h;
  p.stride_w = a->stride_w;
  p.ktot = a->taps_h * a->taps_w * a->Cin;
  p.gw = a->gw;
  p.staged = g_tuning.wgrad_staged;

  // pixel block: 64 pixels unless a stage would not leave room for >= 3 stages
  int P = 64;
  {
    const uint32_t stage64 = 2u * 64 * 128 + (uint32_t)p.G * (p.block_c / 64) * 64 * 128;
    if (stage64 * 3 > 225u * 1024u - (p.staged ? kWgStgBytes : 0u)) P = 32;
  }
  const int npix = a->Ho * a->Wo;
  while (P > 1 && (int64_t)P > (int64_t)npix * a->B) P >>= 1;  // tiny problems
  if (P < 16) P = 16;
  p.p_log2 = ilog2(P);
  const int wo_p2 = 1 << ilog2(a->Wo), ho_p2 = 1 << ilog2(a->Ho);
  const int bw = wo_p2 < P ? wo_p2 : P;
  int bh = P / bw;
  if (bh > ho_p2) bh = ho_p2;
  const int bn = P / (bw * bh);
  p.bw_log2 = ilog2(bw);
  p.bh_log2 = ilog2(bh);
  p.bn_log2 = ilog2(bn);
  p.tiles_w = (a->Wo + bw - 1) / bw;
  p.tiles_h = (a->Ho + bh - 1) / bh;
  p.tiles_b = (a->B + bn - 1) / bn;
  const int k_tiles = p.tiles_w * p.tiles_h * p.tiles_b;

  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int base_items = p.m_tiles * p.groups;
  const int items_override = g_tuning.wgrad_items_per_sm;     // tbg_set_tuning("wgrad_items_per_sm", n); 0 = heuristic
  // tbg_set_tuning("wgrad_items_per_sm", n) forces n rounds; 0 = cost model (host_util.h::pick_splits)
  int splits = items_override > 0 ? (items_override * sms) / base_items : pick_splits(base_items, k_tiles, sms, 8.0);
  if (splits > k_tiles) splits = k_tiles;
  if (splits < 1) splits = 1;
  // every split must own at least one pixel tile
  {
    const int per = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + per - 1) / per;
  }
  p.splits = splits;

  const uint32_t chunk_bytes = (uint32_t)P * 128u;
  const uint32_t stage_bytes = 2u * chunk_bytes + (uint32_t)p.G * (p.block_c / 64) * chunk_bytes;
  const uint32_t stg_bytes = p.staged ? kWgStgBytes : 0u;
  const uint32_t budget = 227u * 1024u - 1024u - 256u - stg_bytes;
  int stages = (int)(budget / stage_bytes);
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  TBG_CHECK_ARG(stages >= 2, "tbg_conv2d_wgrad: stage too large (%u bytes)", stage_bytes);
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + 256 + stg_bytes;

  CUtensorMap tmGY, tmX;
  {
    const int gH = a->up_h ? 2 * a->Ho : a->Ho, gW = a->up_w ? 2 * a->Wo : a->Wo;
    const int esh = a->up_h ? 2 : 1, esw = a->up_w ? 2 : 1;
    const uint64_t dims[4] = {(uint64_t)a->cout, (uint64_t)gW, (uint64_t)gH, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->cout * 2, (uint64_t)gW * a->cout * 2, (uint64_t)gH * gW * a->cout * 2};
    const uint32_t box[4] = {64, (uint32_t)(bw * esw), (uint32_t)(bh * esh), (uint32_t)bn};
    const uint32_t estr[4] = {1, (uint32_t)esw, (uint32_t)esh, 1};
    int rc = encode_tmap_bf16(&tmGY, a->gy, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)(bw * a->stride_w), (uint32_t)(bh * a->stride_h), (uint32_t)bn};
    const uint32_t estr[4] = {1, (uint32_t)a->stride_w, (uint32_t)a->stride_h, 1};
    int rc = encode_tmap_bf16(&tmX, a->x, 4, dims, strides, box, estr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int total_items = p.m_tiles * p.groups * p.splits;
  const int grid = total_items < sms ? total_items : sms;
  conv_wgrad_kernel<<<grid, 256, smem_bytes, stream>>>(tmGY, tmX, p);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
