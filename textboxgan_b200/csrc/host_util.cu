#include "host_util.h"

#include <string.h>

#include <mutex>

namespace tbg {

static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};

char* last_error_buf() { return g_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, const uint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return set_error(TBG_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  cuuint64_t gdims[5];
  cuuint64_t gstr[5];
  cuuint32_t gbox[5];
  cuuint32_t gel[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    gel[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox,
                  gel, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(TBG_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] "
                     "estr=[%u,%u,%u,%u] base=%p",
                     (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)(rank > 1 ? gdims[1] : 0),
                     (unsigned long long)(rank > 2 ? gdims[2] : 0), (unsigned long long)(rank > 3 ? gdims[3] : 0),
                     gbox[0], rank > 1 ? gbox[1] : 0, rank > 2 ? gbox[2] : 0, rank > 3 ? gbox[3] : 0, gel[0],
                     rank > 1 ? gel[1] : 0, rank > 2 ? gel[2] : 0, rank > 3 ? gel[3] : 0, base);
  }
  return 0;
}

}  // namespace tbg

// CRC-32C (Castagnoli) of a host buffer, slicing-by-8 — the checksum of TensorFlow tensor bundles (checkpoint import,
// textboxgan_b200/tf_checkpoint.py).  Host-only helper: no device work.
static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;
static void crc_init() {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0x82F63B78u : 0u);
    g_crc_table[0][i] = c;
  }
  for (int t = 1; t < 8; ++t)
    for (uint32_t i = 0; i < 256; ++i) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  g_crc_ready = true;
}

extern "C" {
unsigned int tbg_crc32c(const void* data, unsigned long long n, unsigned int crc) {
  if (!g_crc_ready) crc_init();
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_table[7][w & 0xFF] ^ g_crc_table[6][(w >> 8) & 0xFF] ^ g_crc_table[5][(w >> 16) & 0xFF] ^
        g_crc_table[4][(w >> 24) & 0xFF] ^ g_crc_table[3][(w >> 32) & 0xFF] ^ g_crc_table[2][(w >> 40) & 0xFF] ^
        g_crc_table[1][(w >> 48) & 0xFF] ^ g_crc_table[0][(w >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n--) c = g_crc_table[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
const char* tbg_last_error(void) { return tbg::last_error_buf(); }
int tbg_version(void) { return 1; }
long long tbg_launch_count(void) { return tbg::g_launches.load(); }
void tbg_reset_launch_count(void) { tbg::g_launches.store(0); }
}
