// EXPERIMENTAL (branch wip/r02-unvalidated-kernels; depends on the outcome of scripts/exp_halo_umma.cu):
// 3x3 stride-1 SAME convolution whose nine taps share ONE activation halo box per 64-channel block.
//
// conv_igemm_kernel re-loads a 16 KB activation box for every tap and a weight box for every M tile; ncu shows
// it bound by the L2 -> SM feed (profiles/r01d_ncu_full_conv_kernels_summary.txt).  Here a work item is a
// 16-wide x 16-high pixel block of one image = two 8 x 16 M tiles (M = 128 rows each, row r = y*8 + x):
//   * one TMA box 64 ch x 24 x 18 pixels (pitch 24 = 16 + 2 halo, rounded up to a multiple of 8 rows so that
//     every 8-row core group of every tap window has the same swizzle phase) per 64-channel block: 54 KB instead of
//     2 x 9 x 16 KB;
//   * the (tap, sub-tile) operand is a WINDOW of that box: descriptor start = box + ((ty*24 + tx + 8*sub) * 128) B,
//     SBO = 24 * 128 B, optional base_offset = (start >> 7) & 7;
//   * each weight box (N x 64 ch, one per tap and channel block) feeds both sub-tiles.
// L2 traffic per 64-channel block and work item: 54 KB + 9 x N*128 B  (N = 128: 198 KB for 2 x 9 x 4 MMAs = 4608
// tensor clocks -> 43 B/clk/SM, versus 125 B/clk/SM in conv_igemm_kernel).
// Restrictions of this first version: 3x3, stride 1, pad 1, Cin % 64 == 0, N = Cout in {32, 64, 128}, W % 16 == 0,
// H % 16 == 0.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct HaloParams {
  int B, H, W;
  int cin_chunks, cin, cout;
  int tiles_w, tiles_h;          // 16 x 16 pixel work items per image
  int use_base_offset;
  const float* col_scale;
  const float* bias;
  const float* noise;
  const float* noise_strength;
  int act;
  float act_gain;
  void* out;                     // bf16 [B, H, W, cout]
};

static constexpr int kHaloPitch = 24, kHaloRows = 18;
static constexpr uint32_t kHaloBytes = kHaloPitch * kHaloRows * 128;   // 55296
static constexpr int kHaloAStages = 2, kHaloBStages = 4;
static constexpr uint32_t kHStgRow = 256 + 16;
static constexpr uint32_t kHStgBytes = 4 * 32 * kHStgRow;

__device__ __forceinline__ uint64_t halo_desc(uint32_t addr, uint32_t sbo, int use_bo) {
  uint64_t d = umma_smem_desc_sw128(addr, 0, sbo);
  if (use_bo) d |= static_cast<uint64_t>((addr >> 7) & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(256, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t b_bytes = static_cast<uint32_t>(p.cout) * 128u;
  uint8_t* smA = smem;                                         // kHaloAStages x 54 KB (each 1024-aligned: 55296 = 54*1024)
  uint8_t* smB = smem + kHaloAStages * kHaloBytes;             // kHaloBStages x N*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kHaloBStages * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kHaloAStages;
  uint64_t* b_full = bars + 2 * kHaloAStages;
  uint64_t* b_empty = b_full + kHaloBStages;
  uint64_t* tfull = b_empty + kHaloBStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* stg_base = reinterpret_cast<uint8_t*>(tempty + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kHaloAStages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kHaloBStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
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
  const int N = p.cout;
  const int total_items = p.tiles_w * p.tiles_h * p.B;

  if (warp == 0) {
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const int tw = item % p.tiles_w, th = (item / p.tiles_w) % p.tiles_h, b = item / (p.tiles_w * p.tiles_h);
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          mbar_arrive_expect_tx(&a_full[as], kHaloBytes);
          tma_load_4d(smA + as * kHaloBytes, &tmA, &a_full[as], ch * 64, tw * 16 - 1, th * 16 - 1, b);
          if (++as == kHaloAStages) { as = 0; aph ^= 1u; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[bs], bph ^ 1u);
            mbar_arrive_expect_tx(&b_full[bs], b_bytes);
            tma_load_2d(smB + bs * b_bytes, &tmB, &b_full[bs], tap * p.cin + ch * 64, 0);
            if (++bs == kHaloBStages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(128, static_cast<uint32_t>(N), 0, 0);
      int as = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
        const int acc_stage = it & 1;
        mbar_wait(&tempty[acc_stage], ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc_stage * 2 * N);
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smA + as * kHaloBytes);
          for (int tap = 0; tap < 9; ++tap) {
            const int ty = tap / 3, tx = tap - ty * 3;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(smB + bs * b_bytes);
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t win = a_base + static_cast<uint32_t>((ty * kHaloPitch + tx + 8 * sub) * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = halo_desc(win + k * 32, kHaloPitch * 128, p.use_base_offset);
                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                umma_bf16(d_tmem + static_cast<uint32_t>(sub * N), da, db, idesc, (ch | tap | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&b_empty[bs]);
            if (++bs == kHaloBStages) { bs = 0; bph ^= 1u; }
          }
          umma_commit(&a_empty[as]);
          if (++as == kHaloAStages) { as = 0; aph ^= 1u; }
        }
        umma_commit(&tfull[acc_stage]);
      }
    }
  } else if (warp >= 4) {
    const int e = warp - 4;
    const int r = e * 32 + lane;
    const int x_in = r & 7, y_in = r >> 3;                    // M row r = y*8 + x of an 8 x 16 sub-tile
    const float nstr = (p.noise != nullptr) ? __ldg(p.noise_strength) : 0.f;
    uint8_t* const stg = stg_base + e * (32 * kHStgRow);
    uint8_t* const my_row = stg + lane * kHStgRow;
    const int chunk_cols = N < 128 ? N : 128;                  // bf16: 256 bytes per row and flush
    const int j_per_chunk = chunk_cols / 32;
    int it = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++it) {
      const int tw = item % p.tiles_w, th = (item / p.tiles_w) % p.tiles_h, b = item / (p.tiles_w * p.tiles_h);
      const int acc_stage = it & 1;
      mbar_wait(&tfull[acc_stage], (it >> 1) & 1);
      tc_fence_after();
      for (int sub = 0; sub < 2; ++sub) {
        const int oy = th * 16 + y_in, ox = tw * 16 + 8 * sub + x_in;
        const size_t pix = (static_cast<size_t>(b) * p.H + oy) * p.W + ox;
        const float nz = (p.noise != nullptr) ? __ldg(p.noise + pix) * nstr : 0.f;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + static_cast<uint32_t>((acc_stage * 2 + sub) * N);
        long long row_off = 0;
        for (int j = 0; j < N / 32; ++j) {
          uint32_t v[32];
          tmem_ld_32x32(t_row + j * 32, v);
          tmem_ld_wait();
          const int c0 = j * 32;
          if (j % j_per_chunk == 0) row_off = static_cast<long long>(pix * p.cout + c0);
          uint8_t* const srow = my_row + (j % j_per_chunk) * 64;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = c0 + g * 8 + i;
              float val = __uint_as_float(v[g * 8 + i]);
              if (p.col_scale) val *= __ldg(p.col_scale + static_cast<size_t>(b) * p.cout + c);
              val += nz;
              if (p.bias) val += __ldg(p.bias + c);
              if (p.act == 1) val = val > 0.f ? val : 0.2f * val;
              f[i] = val * p.act_gain;
            }
            uint4 pk;
            pk.x = pack_bf16x2(f[0], f[1]);
            pk.y = pack_bf16x2(f[2], f[3]);
            pk.z = pack_bf16x2(f[4], f[5]);
            pk.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(srow + g * 16) = pk;
          }
          if ((j + 1) % j_per_chunk == 0) {
            __syncwarp();
            const int lanes_per_row = (chunk_cols * 2) >> 4;
            const int rows_per_pass = 32 / lanes_per_row;
            const int sbl = lane % lanes_per_row;
            for (int r0 = 0; r0 < 32; r0 += rows_per_pass) {
              const int rr = r0 + lane / lanes_per_row;
              const long long o_el = __shfl_sync(0xffffffffu, row_off, rr);
              const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * kHStgRow + sbl * 16);
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + static_cast<size_t>(o_el) * 2 + sbl * 16) = val;
            }
            __syncwarp();
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

}  // namespace tbg

using namespace tbg;

// Experimental entry: same tensors as tbg_conv2d_igemm for a 3x3 SAME conv (include/tbg.h).
extern "C" int tbg_conv3x3_halo(const void* x, const void* w, void* out, int B, int H, int W, int Cin, int Cout,
                                const float* col_scale, const float* bias, const float* noise, const float* noise_strength,
                                int act, float act_gain, int use_base_offset, void* stream_v) {
  TBG_CHECK_ARG(x && w && out, "tbg_conv3x3_halo: null tensor pointer");
  TBG_CHECK_ARG(B >= 1 && H >= 16 && W >= 16 && H % 16 == 0 && W % 16 == 0, "tbg_conv3x3_halo: H, W must be multiples of 16");
  TBG_CHECK_ARG(Cin >= 64 && Cin % 64 == 0, "tbg_conv3x3_halo: Cin must be a multiple of 64");
  TBG_CHECK_ARG(Cout == 32 || Cout == 64 || Cout == 128, "tbg_conv3x3_halo: Cout must be 32, 64 or 128");
  TBG_CHECK_ARG(!noise || noise_strength, "tbg_conv3x3_halo: noise without noise_strength");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  HaloParams p{};
  p.B = B; p.H = H; p.W = W;
  p.cin = Cin; p.cin_chunks = Cin / 64; p.cout = Cout;
  p.tiles_w = W / 16; p.tiles_h = H / 16;
  p.use_base_offset = use_base_offset;
  p.col_scale = col_scale; p.bias = bias; p.noise = noise; p.noise_strength = noise_strength;
  p.act = act; p.act_gain = act_gain; p.out = out;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[4] = {0, (uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)kHaloPitch, (uint32_t)kHaloRows, 1};
    int rc = encode_tmap_bf16(&tmA, x, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t K = 9ull * Cin;
    const uint64_t dims[2] = {K, (uint64_t)Cout};
    const uint64_t strides[2] = {0, K * 2};
    const uint32_t box[2] = {64, (uint32_t)Cout};
    int rc = encode_tmap_bf16(&tmB, w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const size_t smem_bytes = kHaloAStages * kHaloBytes + kHaloBStages * (size_t)Cout * 128 + 256 + kHStgBytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  TBG_CHECK_ARG(smem_bytes <= 227 * 1024, "tbg_conv3x3_halo: shared memory budget exceeded (%zu)", smem_bytes);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total = p.tiles_w * p.tiles_h * B;
  conv3x3_halo_kernel<<<total < sms ? total : sms, 256, smem_bytes, stream>>>(tmA, tmB, p);
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}
