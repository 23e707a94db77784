// 3x3 stride-1 SAME convolution whose nine taps share ONE activation halo box per 64-channel block.
//
// conv_igemm_kernel re-loads a 16 KB activation box for every tap and a weight box for every M tile, and is bound by the
// L2 -> SM operand feed (~52 B/clk/SM delivered, 125 B/clk/SM needed at full tensor rate: profiles/
// r01d_ncu_full_conv_kernels_summary.txt).  Here a work item is a 16-wide x 16-high pixel block of one image = two
// 8 x 16 M tiles (M = 128 rows each, row r = y*8 + x) times one N tile (<= 128 output channels):
//   * one TMA box 64 ch x 24 x 18 pixels (pitch 24 = 16 + 2 halo, rounded up to a multiple of 8 pixels so that every
//     row of the box starts on a 1024-byte swizzle-atom boundary) per 64-channel block: 54 KB instead of 2 x 9 x 16 KB;
//   * the (tap, sub-tile) operand is a WINDOW of that box: descriptor start = box + ((ty*24 + tx + 8*sub) * 128) B,
//     SBO = 24 * 128 B.  tcgen05 applies the 128B swizzle to absolute shared-memory address bits, so a window that
//     starts at any 128-byte row of the TMA-written box reads back the right elements (scripts/exp_halo_umma.cu,
//     measured on B200: all nine taps exact for pitches 10/16/18/24, K-major and MN-major, with base_offset 0);
//   * each weight box (N x 64 ch, one per tap and channel block) feeds both sub-tiles.
// L2 traffic per 64-channel block and work item: 54 KB + 9 x N*128 B  (N = 128: 198 KB for 2 x 9 x 4 MMAs = 4608
// tensor clocks -> 43 B/clk/SM).  Measured (profiles/r02a_perf_halo.log, batch 32): 64x256 128->128 1151 TF/s against
// 554-953 TF/s for conv_igemm_kernel.
//
// Epilogue: eight warps (two per TMEM lane quadrant, one per sub-tile), per-column vectors (demodulation scale, bias)
// broadcast from shared memory, rows transposed through shared memory so that global stores are whole 128-byte lines.
// Up-sampling geometries (columns = (phase_y, phase_x, cout), the input-gradient of a folded stride-2 convolution)
// scatter phase (py,px) of row (b,i,j) to pixel (2i+py, 2j+px) and may skip structurally-zero taps per phase.
//
// Reference behaviour covered: ModulatedConv2D.call (modulated_conv2d.py:66-122, non-fused algebra), Conv2D.call
// (conv.py:51-73), Noise.call (noise.py:12-22), BiasAct.call (bias_act.py:25-34) and their input gradients.
#include "common.cuh"
#include "host_util.h"

namespace tbg {

struct HaloParams {
  int B, H, W;
  int cin_chunks, cin, cout;
  int block_n, tiles_n;          // N tile (32/64/128 columns, inside one output phase) and their number over n_total
  int tiles_w, tiles_h;          // 16 x 16 pixel work items per image
  int up_h, up_w;
  int out_H, out_W;
  uint32_t tap_mask[4];          // per output phase: bit (ty*3+tx) = tap computed
  const float* col_scale;
  const float* bias;
  const float* noise;
  const float* noise_strength;
  int act;
  float act_gain;
  void* out;                     // bf16 [B, out_H, out_W, cout]
  int a_stages, b_stages;        // halo boxes / weight boxes in flight (the kernel is bound by TMA bytes in flight)
  int staged;                    // epilogue stores through the shared-memory transpose (else 16-byte pieces per thread)
};

static constexpr int kHaloPitch = 24, kHaloRows = 18;
static constexpr uint32_t kHaloBytes = kHaloPitch * kHaloRows * 128;   // 55296 = 54 * 1024
static constexpr int kHaloMaxA = 3, kHaloMaxB = 8;
static constexpr int kHaloEpiWarps = 8;
static constexpr uint32_t kHStgRow = 128 + 16;                          // 64 bf16 columns of a row + padding
static constexpr uint32_t kHStgWarp = 32 * kHStgRow;                    // 4608 B per epilogue warp
static constexpr uint32_t kHVecWarp = 2 * 128 * 4;                      // scale + bias vectors per epilogue warp
static constexpr int kHaloThreads = 128 + 32 * kHaloEpiWarps;

// CTA2: two CTAs of a cluster (one TPC) work on two neighbouring pixel blocks with ONE tcgen05.mma.cta_group::2 per
// (tap, sub-tile, K step): M = 256 rows, each CTA stages its own halo box and only HALF of every weight box (rows
// rank*N/2 ..), which removes a quarter of the shared-memory operand reads per SM and halves the weight traffic.  Both
// producers credit the LEADER's full barriers; the leader's commits arrive on the empty / accumulator-full barriers of
// both CTAs; both epilogues release the leader's accumulator-empty barrier.
template <bool CTA2>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int N = p.block_n;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0u;
  const uint32_t b_bytes = static_cast<uint32_t>(CTA2 ? N / 2 : N) * 128u;     // weight box staged by THIS CTA
  const int kHaloAStages = p.a_stages, kHaloBStages = p.b_stages;
  uint8_t* smA = smem;                                         // a_stages x 54 KB (each 1024-aligned)
  uint8_t* smB = smem + kHaloAStages * kHaloBytes;             // b_stages x N*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + kHaloBStages * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kHaloMaxA;
  uint64_t* b_full = bars + 2 * kHaloMaxA;
  uint64_t* b_empty = b_full + kHaloMaxB;
  uint64_t* tfull = b_empty + kHaloMaxB;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  float* vec_base = reinterpret_cast<float*>(tempty + 4);                      // 16-byte aligned
  uint8_t* stg_base = reinterpret_cast<uint8_t*>(vec_base) + kHaloEpiWarps * kHVecWarp;   // staged path only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < kHaloAStages; ++i) {
      mbar_init(&a_full[i], CTA2 ? 2 : 1);                     // pair: one arrive.expect_tx per producer
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < kHaloBStages; ++i) {
      mbar_init(&b_full[i], CTA2 ? 2 : 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], CTA2 ? 2 * kHaloEpiWarps : 32 * kHaloEpiWarps);   // pair: one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) {
      tmem_alloc_pair(tmem_ptr, 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();                                // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_px = p.tiles_w * p.tiles_h * p.B;
  // item = n_tile-major: consecutive items of a CTA differ in pixels.  Pair mode: a pair item is two neighbouring pixel
  // blocks (2k, 2k+1) times an N tile; CTA `rank` owns block 2k + rank.
  const int px_units = CTA2 ? tiles_px / 2 : tiles_px;
  const int total_items = px_units * p.tiles_n;
  const int item0 = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  auto px_of = [&](int item, int n_tile) -> int {
    const int u = item - n_tile * px_units;
    return CTA2 ? 2 * u + static_cast<int>(rank) : u;
  };
  const bool has_up = (p.up_h | p.up_w) != 0;
  auto phase_of = [&](int n_tile) -> int { return has_up ? (n_tile * N) / p.cout : 0; };

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int item = item0; item < total_items; item += item_step) {
        const int n_tile = item / px_units, px_item = px_of(item, n_tile);
        const int tw = px_item % p.tiles_w, th = (px_item / p.tiles_w) % p.tiles_h, b = px_item / (p.tiles_w * p.tiles_h);
        const uint32_t mask = p.tap_mask[phase_of(n_tile) & 3];
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_empty[as], aph ^ 1u);
          if (CTA2) {
            mbar_arrive_expect_tx_leader(&a_full[as], kHaloBytes);
            tma_load_4d_pair(smA + as * kHaloBytes, &tmA, &a_full[as], ch * 64, tw * 16 - 1, th * 16 - 1, b);
          } else {
            mbar_arrive_expect_tx(&a_full[as], kHaloBytes);
            tma_load_4d(smA + as * kHaloBytes, &tmA, &a_full[as], ch * 64, tw * 16 - 1, th * 16 - 1, b);
          }
          if (++as == kHaloAStages) { as = 0; aph ^= 1u; }
          for (int tap = 0; tap < 9; ++tap) {
            if (!((mask >> tap) & 1u)) continue;
            mbar_wait(&b_empty[bs], bph ^ 1u);
            if (CTA2) {
              mbar_arrive_expect_tx_leader(&b_full[bs], b_bytes);
              tma_load_2d_pair(smB + bs * b_bytes, &tmB, &b_full[bs], tap * p.cin + ch * 64,
                               n_tile * N + static_cast<int>(rank) * (N / 2));
            } else {
              mbar_arrive_expect_tx(&b_full[bs], b_bytes);
              tma_load_2d(smB + bs * b_bytes, &tmB, &b_full[bs], tap * p.cin + ch * 64, n_tile * N);
            }
            if (++bs == kHaloBStages) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(CTA2 ? 256 : 128, static_cast<uint32_t>(N), 0, 0);
      int as = 0, bs = 0, it = 0;
      uint32_t aph = 0, bph = 0;
      for (int item = item0; item < total_items; item += item_step, ++it) {
        const int n_tile = item / px_units;
        const uint32_t mask = p.tap_mask[phase_of(n_tile) & 3];
        const int acc_stage = it & 1;
        mbar_wait(&tempty[acc_stage], ((it >> 1) & 1) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc_stage * 2 * N);
        uint32_t accumulate = 0;
        for (int ch = 0; ch < p.cin_chunks; ++ch) {
          mbar_wait(&a_full[as], aph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smA + as * kHaloBytes);
          for (int tap = 0; tap < 9; ++tap) {
            if (!((mask >> tap) & 1u)) continue;
            const int ty = tap / 3, tx = tap - ty * 3;
            mbar_wait(&b_full[bs], bph);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(smB + bs * b_bytes);
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const uint32_t win = a_base + static_cast<uint32_t>((ty * kHaloPitch + tx + 8 * sub) * 128);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_smem_desc_sw128(win + k * 32, 0, kHaloPitch * 128);
                const uint64_t db = umma_smem_desc_sw128(b_addr + k * 32, 0, 1024);
                if (CTA2)
                  umma_bf16_pair(d_tmem + static_cast<uint32_t>(sub * N), da, db, idesc, (accumulate | k) != 0 ? 1u : 0u);
                else
                  umma_bf16(d_tmem + static_cast<uint32_t>(sub * N), da, db, idesc, (accumulate | k) != 0 ? 1u : 0u);
              }
            }
            accumulate = 1;
            if (CTA2) umma_commit_pair(&b_empty[bs]); else umma_commit(&b_empty[bs]);
            if (++bs == kHaloBStages) { bs = 0; bph ^= 1u; }
          }
          if (CTA2) umma_commit_pair(&a_empty[as]); else umma_commit(&a_empty[as]);
          if (++as == kHaloAStages) { as = 0; aph ^= 1u; }
        }
        if (CTA2) umma_commit_pair(&tfull[acc_stage]); else umma_commit(&tfull[acc_stage]);
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int e = warp - 4;
    const int quad = e & 3;                                    // TMEM lane quadrant this warp may read (= warp % 4)
    const int sub = e >> 2;                                    // sub-tile handled by this warp
    const int r = quad * 32 + lane;
    const int x_in = r & 7, y_in = r >> 3;                     // M row r = y*8 + x of an 8 x 16 sub-tile
    const float nstr = (p.noise != nullptr) ? __ldg(p.noise_strength) : 0.f;
    uint8_t* const stg = stg_base + e * kHStgWarp;
    uint8_t* const my_row = stg + lane * kHStgRow;
    float* const vscale = vec_base + e * 256;
    float* const vbias = vscale + 128;
    const int chunk_cols = N < 64 ? N : 64;                    // bf16: up to 128 bytes per row and flush
    const int j_per_chunk = chunk_cols / 32;
    const int lanes_per_row = (chunk_cols * 2) >> 4;           // 4 or 8
    const int rows_per_pass = 32 / lanes_per_row;
    const int sbl = lane % lanes_per_row;
    int it = 0;
    for (int item = item0; item < total_items; item += item_step, ++it) {
      const int n_tile = item / px_units, px_item = px_of(item, n_tile);
      const int tw = px_item % p.tiles_w, th = (px_item / p.tiles_w) % p.tiles_h, b = px_item / (p.tiles_w * p.tiles_h);
      const int acc_stage = it & 1;
      // per-column vectors of this item -> shared memory (broadcast reads below); overlaps the MMAs of the item.
      // With up-sampling an N tile may hold several output phases (cout < N): column -> (phase, channel).
      const int col0 = n_tile * N;
      __syncwarp();
      if (lane * 4 < N) {
        const int c = has_up ? (col0 + lane * 4) % p.cout : col0 + lane * 4;
        float4 sv = make_float4(1.f, 1.f, 1.f, 1.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.col_scale) sv = __ldg(reinterpret_cast<const float4*>(p.col_scale + static_cast<size_t>(b) * p.cout + c));
        if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        reinterpret_cast<float4*>(vscale)[lane] = sv;
        reinterpret_cast<float4*>(vbias)[lane] = bv;
      }
      __syncwarp();
      const int iy = th * 16 + y_in, ix = tw * 16 + 8 * sub + x_in;
      float nz = 0.f;
      long long row_el = 0;                                    // first element of this row's current column chunk
      mbar_wait(&tfull[acc_stage], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>((acc_stage * 2 + sub) * N);
      for (int j = 0; j < N / 32; ++j) {
        if (j % j_per_chunk == 0) {
          // a chunk (<= 64 columns, cout % chunk == 0) lies inside one output phase
          int c_base = col0 + j * 32, py = 0, px = 0;
          if (has_up) {
            const int ph = c_base / p.cout;
            c_base -= ph * p.cout;
            py = p.up_w ? (ph >> 1) : ph;
            px = p.up_w ? (ph & 1) : 0;
          }
          const int oy = p.up_h ? 2 * iy + py : iy, ox = p.up_w ? 2 * ix + px : ix;
          const size_t pix = (static_cast<size_t>(b) * p.out_H + oy) * p.out_W + ox;
          nz = (p.noise != nullptr) ? __ldg(p.noise + pix) * nstr : 0.f;
          row_el = static_cast<long long>(pix * p.cout + c_base);
        }
        uint32_t v[32];
        tmem_ld_32x32(t_row + j * 32, v);
        tmem_ld_wait();
        uint8_t* const srow = my_row + (j % j_per_chunk) * 64;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 s0 = *reinterpret_cast<const float4*>(vscale + j * 32 + g * 8);
          const float4 s1 = *reinterpret_cast<const float4*>(vscale + j * 32 + g * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(vbias + j * 32 + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(vbias + j * 32 + g * 8 + 4);
          float f[8];
          f[0] = fmaf(__uint_as_float(v[g * 8 + 0]), s0.x, nz) + b0.x;
          f[1] = fmaf(__uint_as_float(v[g * 8 + 1]), s0.y, nz) + b0.y;
          f[2] = fmaf(__uint_as_float(v[g * 8 + 2]), s0.z, nz) + b0.z;
          f[3] = fmaf(__uint_as_float(v[g * 8 + 3]), s0.w, nz) + b0.w;
          f[4] = fmaf(__uint_as_float(v[g * 8 + 4]), s1.x, nz) + b1.x;
          f[5] = fmaf(__uint_as_float(v[g * 8 + 5]), s1.y, nz) + b1.y;
          f[6] = fmaf(__uint_as_float(v[g * 8 + 6]), s1.z, nz) + b1.z;
          f[7] = fmaf(__uint_as_float(v[g * 8 + 7]), s1.w, nz) + b1.w;
          if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = (f[i] > 0.f ? f[i] : 0.2f * f[i]);
          } else if (p.act == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] *= p.act_gain;
          uint4 pk;
          pk.x = pack_bf16x2(f[0], f[1]);
          pk.y = pack_bf16x2(f[2], f[3]);
          pk.z = pack_bf16x2(f[4], f[5]);
          pk.w = pack_bf16x2(f[6], f[7]);
          if (p.staged)
            *reinterpret_cast<uint4*>(srow + g * 16) = pk;
          else
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_el + (j % j_per_chunk) * 32 + g * 8) = pk;
        }
        if (p.staged && (j + 1) % j_per_chunk == 0) {
          // flush the staged chunk: lanes_per_row consecutive lanes write one row's contiguous bytes
          __syncwarp();
          for (int r0 = 0; r0 < 32; r0 += rows_per_pass) {
            const int rr = r0 + lane / lanes_per_row;
            const long long o_el = __shfl_sync(0xffffffffu, row_el, rr);
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * kHStgRow + sbl * 16);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.out) + static_cast<size_t>(o_el) * 2 + sbl * 16) = val;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      if (CTA2) {
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty[acc_stage]);
      } else {
        mbar_arrive(&tempty[acc_stage]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();                                // nobody leaves while the peer may still signal / multicast
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// Does the 3x3 halo kernel cover this tbg_conv2d_igemm call?  (Same arguments; the dispatcher in conv_igemm.cu asks.)
bool conv_halo_applicable(const tbg_conv_args* a) {
  if (a->taps_h != 3 || a->taps_w != 3 || a->pad_h != 1 || a->pad_w != 1 || a->stride_h != 1 || a->stride_w != 1) return false;
  if (a->Ho != a->H || a->Wo != a->W || a->H % 16 != 0 || a->W % 16 != 0 || a->Cin % 64 != 0) return false;
  if (a->residual || a->relu_mask || a->out_fp32) return false;
  if (a->cout % 32 != 0) return false;
  return true;
}

int conv_halo_launch(const tbg_conv_args* a, cudaStream_t stream) {
  HaloParams p{};
  p.B = a->B; p.H = a->H; p.W = a->W;
  p.cin = a->Cin; p.cin_chunks = a->Cin / 64; p.cout = a->cout;
  int block_n = 128;
  while (block_n > 32 && (a->cout % block_n) != 0) block_n >>= 1;
  p.tiles_w = a->W / 16; p.tiles_h = a->H / 16;
  p.up_h = a->up_h; p.up_w = a->up_w;
  p.out_H = a->up_h ? 2 * a->H : a->H;
  p.out_W = a->up_w ? 2 * a->W : a->W;
  const int nph = (1 + a->up_h) * (1 + a->up_w);
  for (int i = 0; i < 4; ++i) {
    p.tap_mask[i] = static_cast<uint32_t>((a->tap_mask[i] ? a->tap_mask[i] : 0x1FFull) & 0x1FFull);
    if (i < nph) TBG_CHECK_ARG(p.tap_mask[i] != 0, "tbg_conv2d_igemm(halo): phase %d has no taps", i);
  }
  // 64 output channels per phase: one 128-column tile holds two neighbouring phases (N = 64 MMAs are shared-memory
  // bound at half the tensor rate) when both compute the same taps
  if (nph > 1 && a->cout == 64 && p.tap_mask[0] == p.tap_mask[1] && (nph == 2 || p.tap_mask[2] == p.tap_mask[3])) block_n = 128;
  p.block_n = block_n;
  p.tiles_n = a->n_total / block_n;
  p.col_scale = a->col_scale; p.bias = a->bias; p.noise = a->noise; p.noise_strength = a->noise_strength;
  p.act = a->act; p.act_gain = a->act_gain; p.out = a->out;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // CTA pairs (tcgen05 cta_group::2): 128-column tiles, an even number of pixel blocks, enough pair items for every TPC
  const int tiles_px_all = p.tiles_w * p.tiles_h * a->B;
  const bool cta2 = g_tuning.halo_cta2 && block_n == 128 && tiles_px_all % 2 == 0 && sms % 2 == 0 &&
                    (tiles_px_all / 2) * p.tiles_n >= sms / 2;
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->B};
    const uint64_t strides[4] = {0, (uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2, (uint64_t)a->H * a->W * a->Cin * 2};
    const uint32_t box[4] = {64, (uint32_t)kHaloPitch, (uint32_t)kHaloRows, 1};
    int rc = encode_tmap_bf16(&tmA, a->x, 4, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint64_t K = 9ull * a->Cin;
    const uint64_t dims[2] = {K, (uint64_t)a->n_total};
    const uint64_t strides[2] = {0, K * 2};
    const uint32_t box[2] = {64, (uint32_t)(cta2 ? block_n / 2 : block_n)};       // pair: each CTA stages half the rows
    int rc = encode_tmap_bf16(&tmB, a->w, 2, dims, strides, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  p.staged = g_tuning.halo_staged;
  p.a_stages = g_tuning.halo_a_stages;
  p.b_stages = g_tuning.halo_b_stages;
  const size_t epi_bytes = kHaloEpiWarps * (kHVecWarp + (p.staged ? kHStgWarp : 0u));
  const size_t b_stage = (size_t)(cta2 ? block_n / 2 : block_n) * 128;
  if (cta2) p.b_stages = 2 * p.b_stages > kHaloMaxB ? kHaloMaxB : 2 * p.b_stages;      // half-size boxes: same bytes in flight
  auto smem_need = [&]() {
    return (size_t)p.a_stages * kHaloBytes + (size_t)p.b_stages * b_stage + 512 + epi_bytes + 1024;
  };
  while (smem_need() > 227 * 1024 && p.b_stages > 2) --p.b_stages;
  while (smem_need() > 227 * 1024 && p.a_stages > 2) --p.a_stages;
  const size_t smem_bytes = smem_need();
  static bool attr_set = false;
  if (!attr_set) {
    TBG_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TBG_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  TBG_CHECK_ARG(smem_bytes <= 227 * 1024, "tbg_conv2d_igemm(halo): shared memory budget exceeded (%zu)", smem_bytes);
  if (cta2) {
    // one cluster of two CTAs per TPC; every pair walks the pair items round-robin
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms, 1, 1);
    cfg.blockDim = dim3(kHaloThreads, 1, 1);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TBG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_halo_kernel<true>, tmA, tmB, p));
  } else {
    const int total = tiles_px_all * p.tiles_n;
    conv3x3_halo_kernel<false><<<total < sms ? total : sms, kHaloThreads, smem_bytes, stream>>>(tmA, tmB, p);
  }
  count_launch();
  TBG_CHECK_CUDA(cudaGetLastError());
  return TBG_OK;
}

}  // namespace tbg
