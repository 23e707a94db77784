// Weight gradient of 3x3 stride-1 SAME convolutions with ONE activation halo box per 64-channel block shared by the taps
// of a work item (the weight-gradient counterpart of csrc/conv_halo.cu).
//
//   gw[n, tap, c] += sum_pixels gy[pixel, n] * x[pixel + tap - 1, c]
//
// conv_wgrad_kernel loads one shifted x box per tap: 80 B/clk/SM of L2 -> SM traffic at full tensor rate for a 128 -> 128
// layer (measured 300-490 TF/s, profiles/r02b_layer_perf.log).  Here the K block is an 8 wide x 16 high pixel patch of one
// image (128 pixels, row r = y*8 + x, the same order in both operands):
//   * gy: two 64-channel boxes [128 pixels][64 ch]               (A operand, MN-major: LBO = 16 KB, SBO = 1024)
//   * x : one halo box 64 ch x 10 x 18 pixels per 64-channel block (B operand, MN-major WINDOW of the box: start =
//         box + ((ty + 2k) * 10 + tx) * 128 B for tap (ty,tx) and 16-pixel K step k, SBO = 10 * 128 B, LBO = distance
//         between the channel blocks).  tcgen05 applies the 128B swizzle to absolute shared-memory address bits, so the
//         shifted window reads back what TMA wrote (scripts/experiments/exp_halo_umma.cu part 2, measured on B200).
// A work item = (128 output channels, up to 128 input channels, a group of 3 taps (5 / 4 for 64-channel items), a range of
// K blocks): the accumulators sit in TMEM (<= 512 columns), every x box feeds all taps of the group.  L2 traffic per K
// block: 32 KB (gy) + 2 x 23 KB (x halo) for 3 x 8 x 64 = 1536 tensor clocks = 51 B/clk/SM, most of it L2 hits because the
// items of one pixel range run side by side.  The accumulators are reduced into fp32 gw with
// vector atomics (split-K over CTAs), like conv_wgrad_kernel.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct WgradHaloParams {
  int B, H, W;
  int cin, cout, n_total;
  int block_c;            // input channels per item: 64 or 128
  int c_tiles;            // cin / block_c
  int m_tiles;            // ceil(cout / 128)
  int taps_per_group;     // 512 / block_c, capped at 9
  int groups;             // ceil(9 / taps_per_group)
  int tiles_w, tiles_h;   // 8 x 16 pixel K blocks per image
  int splits;
  int ktot;               // 9 * cin
  float* gw;
};

static constexpr int kWhPitch = 10, kWhRows = 18;
static constexpr uint32_t kWhHaloChunk = ((kWhPitch * kWhRows * 128 + 1023) / 1024) * 1024;   // 23552: 1024-aligned boxes
static constexpr uint32_t kWhGyChunk = 128 * 128;                                                 // [128 px][64 ch]
static constexpr int kWhStages = 2;

__global__ void __launch_bounds__(256, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmGY, const __grid_constant__ CUtensorMap tmX,
                       const WgradHaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int xchunks = p.block_c / 64;
  const uint32_t stage_bytes = 2u * kWhGyChunk + static_cast<uint32_t>(xchunks) * kWhHaloChunk;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWhStages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kWhStages;
  uint64_t* tfull = bars + 2 * kWhStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmGY);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kWhStages; ++i) {
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

  const int k_tiles = p.tiles_w * p.tiles_h * p.B;
  const int base_items = p.m_tiles * p.c_tiles * p.groups;
  const int total_items = base_items * p.splits;
  // item -> (tap group, c_tile, m_tile, K range); the (group, c_tile, m_tile) index runs fastest, so the CTAs that read the
  // SAME pixel range of gy and x for different taps / channel tiles run side by side and share those lines in L2
  // (ncu, profiles/r02h_ncu_conv_summary.txt: with the K range fastest each tap group re-read both tensors from DRAM,
  // 1.29 GB for 0.54 GB of activations)
  auto decode = [&](int item, int& m_tile, int& c_tile, int& t0, int& t1, int& kt0, int& kt1) {
    int rest = item;
    const int group = rest % p.groups;
    rest /= p.groups;
    c_tile = rest % p.c_tiles;
    rest /= p.c_tiles;
    m_tile = rest % p.m_tiles;
    const int split = rest / p.m_tiles;
    t0 = group * p.taps_per_group;
    t1 = min(9, t0 + p.taps_per_group);
    const int per = (k_tiles + p.splits - 1) / p.splits;
    kt0 = split * per;
    kt1 = min(k_tiles, kt0 + per);
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        int m_tile, c_tile, t0, t1, kt0, kt1;
        decode(item, m_tile, c_tile, t0, t1, kt0, kt1);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int tw = kt % p.tiles_w, th = (kt / p.tiles_w) % p.tiles_h, b = kt / (p.tiles_w * p.tiles_h);
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], 2u * kWhGyChunk + static_cast<uint32_t>(xchunks) * (kWhPitch * kWhRows * 128u));
          uint8_t* sa = smem + stage * stage_bytes;
          for (int j = 0; j < 2; ++j) {
            int c0 = m_tile * 128 + j * 64;
            if (c0 >= p.cout) c0 = p.cout;                       // past the last channel: out of bounds -> zero fill
            tma_load_4d(sa + j * kWhGyChunk, &tmGY, &full[stage], c0, tw * 8, th * 16, b);
          }
          for (int jc = 0; jc < xchunks; ++jc)
            tma_load_4d(sa + 2 * kWhGyChunk + jc * kWhHaloChunk, &tmX, &full[stage], c_tile * p.block_c + jc * 64,
                        tw * 8 - 1, th * 16 - 1, b);
          if (++stage == kWhStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(p.block_c), 1, 1);     // both operands MN-major
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        int m_tile, c_tile, t0, t1, kt0, kt1;
        decode(item, m_tile, c_tile, t0, t1, kt0, kt1);
        mbar_wait(tempty, (it & 1) ^ 1u);
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint32_t x_addr = a_addr + 2u * kWhGyChunk;
          for (int tap = t0; tap < t1; ++tap) {
            const int ty = tap / 3, tx = tap - ty * 3;
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((tap - t0) * p.block_c);
#pragma unroll
            for (int k = 0; k < 8; ++k) {                       // 16 pixels per step = two 8-pixel rows of the patch
              const uint64_t da = umma_smem_desc_sw128(a_addr + k * 2048, kWhGyChunk, 1024);
              const uint64_t db = umma_smem_desc_sw128(x_addr + static_cast<uint32_t>(((ty + 2 * k) * kWhPitch + tx) * 128),
                                                       kWhHaloChunk, kWhPitch * 128);
              umma_bf16(d_tmem, da, db, idesc, (kt > kt0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
          if (++stage == kWhStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull);
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue: TMEM -> red.global.add ================================
    const int e = warp - 4;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      int m_tile, c_tile, t0, t1, kt0, kt1;
      decode(item, m_tile, c_tile, t0, t1, kt0, kt1);
      mbar_wait(tfull, it & 1);
      tc_fence_after();
      const int nn = m_tile * 128 + e * 32 + lane;               // output channel of this thread's accumulator row
      for (int tap = t0; tap < t1; ++tap) {
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + static_cast<uint32_t>((tap - t0) * p.block_c);
        for (int j = 0; j < p.block_c / 32; ++j) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + j * 32, v);
          tmem_ld_wait();
          if (nn < p.n_total && kt1 > kt0) {
            float* dst = p.gw + static_cast<size_t>(nn) * p.ktot + tap * p.cin + c_tile * p.block_c + j * 32;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + g * 4),
                           "f"(__uint_as_float(v[g * 4])), "f"(__uint_as_float(v[g * 4 + 1])),
                           "f"(__uint_as_float(v[g * 4 + 2])), "f"(__uint_as_float(v[g * 4 + 3]))
                           : "memory");
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

bool wgrad_halo_applicable(const tbg_wgrad_args* a) {
  if (a->taps_h != 3 || a->taps_w != 3 || a->pad_h != 1 || a->pad_w != 1 || a->stride_h != 1 || a->stride_w != 1) return false;
  if (a->up_h || a->up_w || a->Ho != a->H || a->Wo != a->W) return false;
  if (a->H % 16 != 0 || a->W % 8 != 0 || a->Cin % 64 != 0 || a->cout % 32 != 0) return false;
  // enough K blocks for the split-K items to amortise their accumulator drain
  return static_cast<long long>(a->B) * (a->H / 16) * (a->W / 8) >= 256;
}

int wgrad_halo_launch(const tbg_wgrad_args* a, cudaStream_t stream) {
  WgradHaloParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.cin = a->Cin; p.cout = a->cout; p.n_total = a->n_total;
  p.block_c = (a->Cin % 128 == 0) ? 128 : 64;
  p.c_tiles = a->Cin / p.block_c;
  p.m_tiles = (a->cout + 127) / 128;
  // tap groups of equal size (3 x 3 taps for 128-channel items, 5 + 4 for 64-channel items): with the static round-robin
  // over CTAs unequal groups (4 + 4 + 1) left the CTAs holding the short items idle (profiles/r02j_wgrad.log)
  int max_taps = 512 / p.block_c;
  if (max_taps > 9) max_taps = 9;
  p.groups = (9 + max_taps - 1) / max_taps;
  p.taps_per_group = (9 + p.groups - 1) / p.groups;
  p.tiles_w = a->W / 8; p.tiles_h = a->H / 16;
  p.ktot = 9 * a->Cin;
  p.gw = a->gw;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int k_tiles = p.tiles_w * p.tiles_h * a->B;
  const int base_items = p.m_tiles * p.c_tiles * p.groups;
  int splits = pick_splits(base_items, k_tiles, sms, 8.0);
  if (splits > k_tiles) splits = k_tiles;
  if (splits < 1) splits = 1;
  {
    const int per = (k_tiles + splits - 1) / splits;
    splits = (k_tiles + per - 1) / per;
  }
  p.splits = splits;

  CUtensorMap tmGY, tmX;
  {
    const uint64_t dims[4] = {(uint64_t)a->cout, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->cout * 2, (uint64_t)a->W * a->cout * 2, (uint64_t)a->H * a->W * a->cout * 2};
    const uint32_t box[4] = {64, 8, 16, 1};
    int rc = encode_tmap_bf16(&tmGY, a->gy, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)kWhPitch, (uint32_t)kWhRows, 1};
    int rc = encode_tmap_bf16(&tmX, a->x, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const size_t smem_bytes = kWhStages * (2u * kWhGyChunk + (size_t)(p.block_c / 64) * kWhHaloChunk) + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int total_items = base_items * p.splits;
  conv_wgrad_halo_kernel<<<total_items < sms ? total_items : sms, 256, smem_bytes, stream>>>(tmGY, tmX, p);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

}  // namespace tbg
