// Host-side helpers of the C ABI: thread-local error string, launch counter, TMA descriptor
// encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/tbg.h"

namespace tbg {

char* last_error_buf();  // thread-local, 512 bytes
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define TBG_CHECK_ARG(cond, ...)                                  \
  do {                                                            \
    if (!(cond)) return tbg::set_error(TBG_ERR_INVALID_ARG, __VA_ARGS__); \
  } while (0)

#define TBG_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return tbg::set_error(TBG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                               \
  } while (0)

// Encode a tiled bf16 tensor map. dims/strides innermost-first; strides in BYTES for dims 1..rank-1.
// Returns 0 or sets the error string.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle);

// Process-wide tuning switches, set through tbg_set_tuning (never from the environment).
// Split-K factor of the weight-gradient kernels: `base_items` independent (tile, tap-group) items are each cut into
// `splits` ranges of K blocks and dealt round-robin to one persistent CTA per SM.  Cost model: every round costs the K
// blocks of one item plus its accumulator drain (fp32 vector atomics of up to 512 TMEM columns, worth ~`drain_k` K blocks;
// single-buffered TMEM, so it overlaps nothing): minimise  rounds(s) * (k_tiles / s + drain_k).  A plain
// "2 items per SM, rounded up" left 2 * sms + 1 items = a whole third round on ONE SM while the others idled
// (ncu: sm__cycles_active 70 % of elapsed, profiles/r02ab_ncu_conv_summary.txt).
inline int pick_splits(int base_items, int k_tiles, int sms, double drain_k) {
  int smax = (4 * sms + base_items - 1) / base_items;
  if (smax > k_tiles) smax = k_tiles;
  if (smax < 1) smax = 1;
  int best = 1;
  double best_t = 1e300;
  for (int s = 1; s <= smax; ++s) {
    const long long items = static_cast<long long>(base_items) * s;
    const long long rounds = (items + sms - 1) / sms;
    const double per = static_cast<double>((k_tiles + s - 1) / s);
    const double t = static_cast<double>(rounds) * (per + drain_k);
    if (t < best_t - 1e-9) {
      best_t = t;
      best = s;
    }
  }
  return best;
}

struct Tuning {
  int conv_halo = 1;           // 3x3 stride-1 convolutions on the halo-reuse kernel when it applies
  int wgrad_staged = 0;        // conv_wgrad: staged vector atomics (measured slower on B200, profiles/r02b_layer_perf.log: off)
  int wgrad_items_per_sm = 0;  // conv_wgrad: split-K work items per SM (0 = heuristic)
  int halo_a_stages = 2;       // conv3x3_halo: activation halo boxes in flight (54 KB each)
  int halo_b_stages = 4;       // conv3x3_halo: weight boxes in flight (N x 128 B each)
  int halo_staged = 1;         // conv3x3_halo epilogue: stores transposed through shared memory
  int wgrad_halo = 1;          // 3x3 stride-1 weight gradients on the halo-reuse kernel when it applies
  int halo_cta2 = 0;           // conv3x3_halo: CTA pairs (tcgen05 cta_group::2) on 128-column tiles
  int lstm_cluster = 1;        // LSTM whole-sequence kernels on 4-CTA clusters with smem-resident W_hh
};
extern Tuning g_tuning;

bool wgrad_halo_applicable(const ::tbg_wgrad_args* a);
int wgrad_halo_launch(const ::tbg_wgrad_args* a, cudaStream_t stream);
bool conv_halo_applicable(const ::tbg_conv_args* a);
int conv_halo_launch(const ::tbg_conv_args* a, cudaStream_t stream);

inline int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace tbg
