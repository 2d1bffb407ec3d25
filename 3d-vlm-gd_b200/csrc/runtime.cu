// lib3dgd runtime: error state, driver entry points (TMA descriptor encoding), debug GEMM entry.
#include "../../include/gd3.h"
#include "common.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm_2sm.cuh"   // experimental 2-CTA kernel: debug entry only

#include <cstdlib>
#include <cstring>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler injects the library

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace gd3 {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfRec {
  const char* name;
  cudaEvent_t a, b;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
}  // namespace

// NVTX: every kernel launch of the library sits inside a named range (the same names gd3_profile_read reports), so
// an nsys / ncu timeline of a training step shows which loss stage a kernel belongs to.  Off unless GD3_NVTX=1:
// the push / pop pair costs ~100 ns per launch even without a profiler attached.
static bool nvtx_on() {
  static const bool on = [] {
    const char* e = getenv("GD3_NVTX");
    return e && e[0] && e[0] != '0';
  }();
  return on;
}

ProfScope::ProfScope(const char* n, cudaStream_t s) : name(n), stream(s), slot(-1) {
  if (nvtx_on()) nvtxRangePushA(n);
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{n, nullptr, nullptr};
  if (!g_prof_pool.empty()) {
    r.a = g_prof_pool.back().first;
    r.b = g_prof_pool.back().second;
    g_prof_pool.pop_back();
  } else if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) {
    return;
  }
  cudaEventRecord(r.a, s);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (nvtx_on()) nvtxRangePop();
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].b, stream);
}

namespace tc {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    // resolved through the runtime so the library does not link libcuda directly
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int64_t k_extent, int64_t rows, int64_t batch,
                   int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GD3_ERR_CUDA;
  }
  GD3_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  GD3_REQUIRE(row_stride_elems % 8 == 0 && (batch <= 1 || batch_stride_elems % 8 == 0),
              "TMA strides must be multiples of 16 bytes (row stride %lld, batch stride %lld elements)",
              (long long)row_stride_elems, (long long)batch_stride_elems);
  GD3_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2,
                           (cuuint64_t)(batch > 1 ? batch_stride_elems : row_stride_elems * rows) * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (k=%lld rows=%lld batch=%lld ld=%lld)", (int)r,
              (long long)k_extent, (long long)rows, (long long)batch, (long long)row_stride_elems);
    return GD3_ERR_CUDA;
  }
  return GD3_OK;
}

// the same view with 32-element (64-byte) boxes and the 64-byte swizzle: half-size pipeline stages for kernels whose
// shared memory only fits two 64-element stages (ap_fused_kernel)
int make_tmap_bf16_k32(CUtensorMap* out, const void* base, int64_t k_extent, int64_t rows, int64_t batch,
                       int64_t row_stride_elems, int64_t batch_stride_elems, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GD3_ERR_CUDA;
  }
  GD3_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  GD3_REQUIRE(row_stride_elems % 8 == 0 && (batch <= 1 || batch_stride_elems % 8 == 0),
              "TMA strides must be multiples of 16 bytes (row stride %lld, batch stride %lld elements)",
              (long long)row_stride_elems, (long long)batch_stride_elems);
  GD3_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
  cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2,
                           (cuuint64_t)(batch > 1 ? batch_stride_elems : row_stride_elems * rows) * 2};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (k32) failed with CUresult %d (k=%lld rows=%lld batch=%lld ld=%lld)", (int)r,
              (long long)k_extent, (long long)rows, (long long)batch, (long long)row_stride_elems);
    return GD3_ERR_CUDA;
  }
  return GD3_OK;
}

int make_tmap_store16(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t batch,
                      int64_t row_stride_elems, int64_t batch_stride_elems, bool fp16) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GD3_ERR_CUDA;
  }
  GD3_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA store base pointer must be 16-byte aligned");
  GD3_REQUIRE(row_stride_elems % 8 == 0 && (batch <= 1 || batch_stride_elems % 8 == 0),
              "TMA store strides must be multiples of 16 bytes (row stride %lld, batch stride %lld elements)",
              (long long)row_stride_elems, (long long)batch_stride_elems);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 2,
                           (cuuint64_t)(batch > 1 ? batch_stride_elems : row_stride_elems * rows) * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (store map) failed with CUresult %d (cols=%lld rows=%lld batch=%lld ld=%lld)", (int)r,
              (long long)cols, (long long)rows, (long long)batch, (long long)row_stride_elems);
    return GD3_ERR_CUDA;
  }
  return GD3_OK;
}

int make_tmap_store32(CUtensorMap* out, const void* base, int64_t cols, int64_t rows, int64_t batch,
                      int64_t row_stride_elems, int64_t batch_stride_elems) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return GD3_ERR_CUDA;
  }
  GD3_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA store base pointer must be 16-byte aligned");
  GD3_REQUIRE(row_stride_elems % 4 == 0 && (batch <= 1 || batch_stride_elems % 4 == 0),
              "TMA store strides must be multiples of 16 bytes (row stride %lld, batch stride %lld fp32 elements)",
              (long long)row_stride_elems, (long long)batch_stride_elems);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride_elems * 4,
                           (cuuint64_t)(batch > 1 ? batch_stride_elems : row_stride_elems * rows) * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 store map) failed with CUresult %d (cols=%lld rows=%lld batch=%lld ld=%lld)",
              (int)r, (long long)cols, (long long)rows, (long long)batch, (long long)row_stride_elems);
    return GD3_ERR_CUDA;
  }
  return GD3_OK;
}

}  // namespace tc
}  // namespace gd3

using namespace gd3;

extern "C" {

int gd3_version(void) { return GD3_VERSION; }

long long gd3_launch_count(void) { return gd3::g_launches.load(); }

void gd3_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(gd3::g_prof_mu);
  gd3::g_prof_on = on != 0;
}

// Writes one line per kernel name: "<name> <launches> <total_ms>\n".  Synchronises the device, then
// (when a buffer is given) clears the records.  Returns the number of bytes needed (excluding the terminator).
size_t gd3_profile_read(char* buf, size_t buf_bytes) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(gd3::g_prof_mu);
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : gd3::g_prof) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
  }
  std::string out;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (buf && buf_bytes) {
    const size_t n = out.size() < buf_bytes - 1 ? out.size() : buf_bytes - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
    for (auto& r : gd3::g_prof) gd3::g_prof_pool.emplace_back(r.a, r.b);   // recycle the events
    gd3::g_prof.clear();
  }
  return out.size();
}
const char* gd3_last_error(void) { return gd3::last_error(); }

int gd3_debug_gemm_bf16(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                        int64_t lda, int64_t ldb, int64_t ldc, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(A && B && C, "gd3_debug_gemm_bf16: null pointer");
  const int bn = tile_n < 0 ? -tile_n : tile_n;
  GD3_REQUIRE(bn == 128 || bn == 192 || bn == 256, "gd3_debug_gemm_bf16: tile_n must be +-128, +-192 or +-256");
  CUtensorMap ta, tb;
  int rc;
  tc::EpiStoreF32::Params ep{C, (int)M, (int)N, ldc, M * ldc, 1.0f, nullptr};
  tc::GemmShape s{(int)M, (int)N, (int)K, (int)batch};
  if (tile_n > 0 && (rc = tc::enable_tma_store(ep, (int)batch))) return rc;
  if ((rc = tc::make_tmap_bf16(&ta, A, K, M, batch, lda, M * lda, tc::BM))) return rc;
  if (tile_n < 0) {
    // 2-CTA (cta_group::2) kernel: 256 x bn tiles, each CTA loads half of the B tile
    if ((rc = tc::make_tmap_bf16(&tb, B, K, N, batch, ldb, N * ldb, bn / 2))) return rc;
    if (bn == 256) return tc::launch_gemm_2sm<256, 8, tc::EpiStoreF32>("debug_gemm_2sm", ta, tb, s, ep, stream);
    if (bn == 192) return tc::launch_gemm_2sm<192, 8, tc::EpiStoreF32>("debug_gemm_2sm", ta, tb, s, ep, stream);
    return tc::launch_gemm_2sm<128, 8, tc::EpiStoreF32>("debug_gemm_2sm", ta, tb, s, ep, stream);
  }
  if ((rc = tc::make_tmap_bf16(&tb, B, K, N, batch, ldb, N * ldb, bn))) return rc;
  if (bn == 256) return tc::launch_gemm<256, 8, tc::EpiStoreF32>("debug_gemm", ta, tb, s, ep, stream);
  if (bn == 192) return tc::launch_gemm<192, 8, tc::EpiStoreF32>("debug_gemm", ta, tb, s, ep, stream);
  return tc::launch_gemm<128, 8, tc::EpiStoreF32>("debug_gemm", ta, tb, s, ep, stream);
}

// Same GEMM with MN-major operands: a_mn -> A is given as (batch, K, M) (M contiguous), b_mn -> B as (batch, K, N).
int gd3_debug_gemm_bf16_mn(const void* A, const void* B, float* C, int64_t M, int64_t N, int64_t K, int64_t batch,
                           int a_mn, int b_mn, int tile_n, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GD3_REQUIRE(A && B && C, "gd3_debug_gemm_bf16_mn: null pointer");
  GD3_REQUIRE(tile_n == 128 || tile_n == 192 || tile_n == 256, "gd3_debug_gemm_bf16_mn: tile_n must be 128, 192 or 256");
  GD3_REQUIRE((!a_mn || M % 8 == 0) && (!b_mn || N % 8 == 0) && ((a_mn && b_mn) || K % 8 == 0),
              "gd3_debug_gemm_bf16_mn: contiguous dimensions must be multiples of 8 (16-byte TMA strides)");
  CUtensorMap ta, tb;
  int rc;
  tc::EpiStoreF32::Params ep{C, (int)M, (int)N, N, M * N, 1.0f, nullptr};
  tc::GemmShape s{(int)M, (int)N, (int)K, (int)batch};
  if (a_mn) rc = tc::make_tmap_bf16(&ta, A, M, K, batch, M, K * M, 64);
  else rc = tc::make_tmap_bf16(&ta, A, K, M, batch, K, M * K, tc::BM);
  if (rc) return rc;
  if (b_mn) rc = tc::make_tmap_bf16(&tb, B, N, K, batch, N, K * N, 64);
  else rc = tc::make_tmap_bf16(&tb, B, K, N, batch, K, N * K, tile_n);
  if (rc) return rc;
#define GD3_DBG_MN(BN)                                                                                                  \
  do {                                                                                                                  \
    if (a_mn && b_mn) return tc::launch_gemm<BN, 8, tc::EpiStoreF32, true, true>("debug_gemm_mn", ta, tb, s, ep, stream); \
    if (a_mn) return tc::launch_gemm<BN, 8, tc::EpiStoreF32, true, false>("debug_gemm_mn", ta, tb, s, ep, stream);       \
    if (b_mn) return tc::launch_gemm<BN, 8, tc::EpiStoreF32, false, true>("debug_gemm_mn", ta, tb, s, ep, stream);       \
    return tc::launch_gemm<BN, 8, tc::EpiStoreF32>("debug_gemm_mn", ta, tb, s, ep, stream);                              \
  } while (0)
  if (tile_n == 256) GD3_DBG_MN(256);
  if (tile_n == 192) GD3_DBG_MN(192);
  GD3_DBG_MN(128);
#undef GD3_DBG_MN
}

}  // extern "C"
