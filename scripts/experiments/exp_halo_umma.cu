// EXPERIMENT (round 2, run first): can tcgen05.mma read SHIFTED WINDOWS of one TMA-loaded halo tile?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I textboxgan_b200/csrc -o gpurun_out/exp_halo \
//        scripts/exp_halo_umma.cu textboxgan_b200/csrc/host_util.cu && gpurun_out/exp_halo
// If yes, the 3x3 taps of the implicit-GEMM convolutions can share one activation box per 64-channel block
// (A traffic / ~5 in conv_igemm, x traffic / ~6 in conv_wgrad) — DESIGN.md section 10.
//
// Part 1 (K-major A, conv_igemm): halo tile [HB rows][pitch px][64 ch] in SW128 smem; the M = 128 rows of a
//   tap's window are an 8-wide x 16-high pixel tile: row r = y*8 + x lives at start + y*SBO + x*128 with
//   start = base + (ty*pitch + tx)*128, SBO = pitch*128.
// Part 2 (MN-major B, conv_wgrad): K = pixels; a K=16 step is 16 consecutive pixels of one row of the halo tile,
//   start = base + ((y+ty)*pitch + tx)*128, SBO = 1024.
// Variants: descriptor base_offset field 0 vs ((start >> 7) & 7); pitch a multiple of 8 vs not.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "host_util.h"

using namespace tbg;

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return umma_smem_desc_sw128(addr, lbo, sbo) | (static_cast<uint64_t>(base_off & 7u) << 49);
}

// ---------------------------------------------------------------------------------------------------------
// Part 1: D[128 x 64] = A_window(ty,tx)[128 x 64] * W[64 x 64]^T for the nine taps, K-major operands.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
halo_kmajor_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, float* out,
                   int pitch, int rows_total, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smX = smem;                     // rows_total * 128 B
  uint8_t* smW = smem + 48 * 1024;         // 8 KB
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 60 * 1024);
  uint64_t* bar_mma = bar_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_full, rows_total * 128 + 64 * 128);
    tma_load_4d(smX, &tmX, bar_full, 0, 0, 0, 0);
    tma_load_2d(smW, &tmW, bar_full, 0, 0);
  }
  mbar_wait(bar_full, 0);
  tc_fence_after();
  const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
  for (int tap = 0; tap < 9; ++tap) {
    const int ty = tap / 3, tx = tap % 3;
    if (threadIdx.x == 0) {
      const uint32_t start = smem_u32(smX) + (ty * pitch + tx) * 128;
      const uint32_t bo = use_base_offset ? ((start >> 7) & 7u) : 0u;
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = desc_sw128_bo(start + k * 32, 0, pitch * 128, bo);
        const uint64_t db = umma_smem_desc_sw128(smem_u32(smW) + k * 32, 0, 1024);
        umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, tap & 1);
    tc_fence_after();
    for (int j = 0; j < 2; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(tap * 128 + warp * 32 + lane) * 64 + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------------------
// Part 2: D[128 x 64] = sum_{64 pixels} gy[pixel][128]^T * xwin(ty,tx)[pixel][64], MN-major operands (wgrad).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
halo_mnmajor_kernel(const __grid_constant__ CUtensorMap tmGY, const __grid_constant__ CUtensorMap tmX, float* out,
                    int pitch, int rows_total, int use_base_offset) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smG = smem;                     // 2 boxes x 64 px x 128 B = 16 KB
  uint8_t* smX = smem + 16 * 1024;         // rows_total * 128 B
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 60 * 1024);
  uint64_t* bar_mma = bar_full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_full, 2 * 64 * 128 + rows_total * 128);
    tma_load_4d(smG, &tmGY, bar_full, 0, 0, 0, 0);
    tma_load_4d(smG + 8192, &tmGY, bar_full, 64, 0, 0, 0);
    tma_load_4d(smX, &tmX, bar_full, 0, 0, 0, 0);
  }
  mbar_wait(bar_full, 0);
  tc_fence_after();
  const uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
  for (int tap = 0; tap < 9; ++tap) {
    const int ty = tap / 3, tx = tap % 3;
    if (threadIdx.x == 0) {
      for (int k = 0; k < 4; ++k) {   // K step = the 16 pixels of tile row k
        const uint64_t da = umma_smem_desc_sw128(smem_u32(smG) + k * 2048, 8192, 1024);
        const uint32_t start = smem_u32(smX) + ((k + ty) * pitch + tx) * 128;
        const uint32_t bo = use_base_offset ? ((start >> 7) & 7u) : 0u;
        const uint64_t db = desc_sw128_bo(start, 8192, 1024, bo);
        umma_bf16(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, tap & 1);
    tc_fence_after();
    for (int j = 0; j < 2; ++j) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j * 32, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) out[(tap * 128 + warp * 32 + lane) * 64 + j * 32 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 1) tmem_dealloc(tmem_base, 64);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main() {
  srand(1);
  auto rnd = []() { return bf((rand() % 2001 - 1000) / 1000.f); };
  cudaFuncSetAttribute(halo_kmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(halo_mnmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  float* d_out;
  CK(cudaMalloc(&d_out, 9 * 128 * 64 * 4));
  std::vector<float> out(9 * 128 * 64);
  // ---------------- part 1 ----------------
  for (int pitch : {16, 10}) {
    const int HB = 18;
    std::vector<float> x(HB * pitch * 64), w(64 * 64);
    for (auto& v : x) v = rnd();
    for (auto& v : w) v = rnd();
    std::vector<__nv_bfloat16> xb(x.size()), wb(w.size());
    for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16(x[i]);
    for (size_t i = 0; i < w.size(); ++i) wb[i] = __float2bfloat16(w[i]);
    __nv_bfloat16 *dx, *dw;
    CK(cudaMalloc(&dx, xb.size() * 2));
    CK(cudaMalloc(&dw, wb.size() * 2));
    CK(cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmX, tmW;
    {
      const uint64_t dims[4] = {64, (uint64_t)pitch, (uint64_t)HB, 1};
      const uint64_t str[4] = {0, 128, (uint64_t)pitch * 128, (uint64_t)HB * pitch * 128};
      const uint32_t box[4] = {64, (uint32_t)pitch, (uint32_t)HB, 1};
      if (encode_tmap_bf16(&tmX, dx, 4, dims, str, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s\n", tbg_last_error()); return 1; }
      const uint64_t d2[2] = {64, 64};
      const uint64_t s2[2] = {0, 128};
      const uint32_t b2[2] = {64, 64};
      if (encode_tmap_bf16(&tmW, dw, 2, d2, s2, b2, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s\n", tbg_last_error()); return 1; }
    }
    for (int bo = 0; bo < 2; ++bo) {
      CK(cudaMemset(d_out, 0, out.size() * 4));
      halo_kmajor_kernel<<<1, 128, 64 * 1024>>>(tmX, tmW, d_out, pitch, HB * pitch, bo);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
      printf("part1 K-major  pitch=%2d base_offset=%s:", pitch, bo ? "start" : "0    ");
      for (int tap = 0; tap < 9; ++tap) {
        const int ty = tap / 3, tx = tap % 3;
        float err = 0.f;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < 64; ++n) {
            const int y = r / 8, xx = r % 8;
            float acc = 0.f;
            for (int c = 0; c < 64; ++c) acc += x[((y + ty) * pitch + xx + tx) * 64 + c] * w[n * 64 + c];
            err = fmaxf(err, fabsf(acc - out[(tap * 128 + r) * 64 + n]));
          }
        printf(" t%d:%.3g", tap, err);
      }
      printf("\n");
    }
    cudaFree(dx);
    cudaFree(dw);
  }
  // ---------------- part 2 ----------------
  for (int pitch : {24, 18}) {
    const int HB = 6;
    std::vector<float> x(HB * pitch * 64), gy(64 * 128);
    for (auto& v : x) v = rnd();
    for (auto& v : gy) v = rnd();
    std::vector<__nv_bfloat16> xb(x.size()), gb(gy.size());
    for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16(x[i]);
    for (size_t i = 0; i < gy.size(); ++i) gb[i] = __float2bfloat16(gy[i]);
    __nv_bfloat16 *dx, *dg;
    CK(cudaMalloc(&dx, xb.size() * 2));
    CK(cudaMalloc(&dg, gb.size() * 2));
    CK(cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg, gb.data(), gb.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmX, tmG;
    {
      const uint64_t dims[4] = {64, (uint64_t)pitch, (uint64_t)HB, 1};
      const uint64_t str[4] = {0, 128, (uint64_t)pitch * 128, (uint64_t)HB * pitch * 128};
      const uint32_t box[4] = {64, (uint32_t)pitch, (uint32_t)HB, 1};
      if (encode_tmap_bf16(&tmX, dx, 4, dims, str, box, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s\n", tbg_last_error()); return 1; }
      const uint64_t dg4[4] = {128, 16, 4, 1};             // gy: [4 rows][16 px][128 ch]
      const uint64_t sg4[4] = {0, 256, 16 * 256, 64 * 256};
      const uint32_t bg4[4] = {64, 16, 4, 1};
      if (encode_tmap_bf16(&tmG, dg, 4, dg4, sg4, bg4, nullptr, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("%s\n", tbg_last_error()); return 1; }
    }
    for (int bo = 0; bo < 2; ++bo) {
      CK(cudaMemset(d_out, 0, out.size() * 4));
      halo_mnmajor_kernel<<<1, 128, 64 * 1024>>>(tmG, tmX, d_out, pitch, HB * pitch, bo);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
      printf("part2 MN-major pitch=%2d base_offset=%s:", pitch, bo ? "start" : "0    ");
      for (int tap = 0; tap < 9; ++tap) {
        const int ty = tap / 3, tx = tap % 3;
        float err = 0.f;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float acc = 0.f;
            for (int y = 0; y < 4; ++y)
              for (int xx = 0; xx < 16; ++xx) acc += gy[(y * 16 + xx) * 128 + m] * x[((y + ty) * pitch + xx + tx) * 64 + n];
            err = fmaxf(err, fabsf(acc - out[(tap * 128 + m) * 64 + n]));
          }
        printf(" t%d:%.3g", tap, err);
      }
      printf("\n");
    }
    cudaFree(dx);
    cudaFree(dg);
  }
  printf("(errors ~1e-5 or below on all nine taps of a line = that addressing variant works)\n");
  return 0;
}
