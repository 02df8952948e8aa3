// Error channel, launch counter and the TMA-descriptor cache of librecformer_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "rf_common.h"

namespace rf {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(RF_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return RF_OK;
}

static nonce_loader_fn g_nonce_loaders[16];
static int g_n_nonce_loaders = 0;
int register_nonce_loader(nonce_loader_fn fn) {
  if (g_n_nonce_loaders < 16) g_nonce_loaders[g_n_nonce_loaders++] = fn;
  return g_n_nonce_loaders;
}

// cudaFuncSetAttribute is per DEVICE: true exactly once per (flag, current device)
bool first_use_on_device(std::atomic<unsigned long long>* seen) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  return (seen->fetch_or(bit) & bit) == 0;
}

static thread_local RowActivity g_activity{nullptr, 0, nullptr, nullptr};
const RowActivity& row_activity() { return g_activity; }
void set_row_activity(const RowActivity& a) { g_activity = a; }

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- TMA descriptors -------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, d2, s1, s2;
  uint32_t b0, b1;
  int dev;
  bool operator==(const MapKey& o) const { return std::memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};

static std::mutex g_map_mu;
// values are heap-allocated so the returned pointers stay valid across rehashes
static std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> g_maps;

static const CUtensorMap* lookup_or_encode(MapKey key, int rank) {
  static_assert(sizeof(MapKey) % 8 == 0, "MapKey must be 8-byte granular");
  cudaGetDevice(&key.dev);
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error(RF_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return nullptr;
  }
  if (g_maps.size() > 8192) {  // unbounded growth guard for callers that keep reallocating
    // Another thread may still hold a pointer it obtained before this eviction (the lock is released before
    // the launch that copies the map into kernel parameters): evicted maps are RETIRED, never freed
    // (128 B each, at most one generation of 8192 per eviction).
    static std::vector<CUtensorMap*> retired;
    for (auto& kv : g_maps) retired.push_back(kv.second);
    g_maps.clear();
  }
  CUtensorMap* m = new CUtensorMap;
  cuuint64_t dims[3] = {key.d0, key.d1, key.d2};
  cuuint64_t strides[2] = {key.s1 * 2, key.s2 * 2};  // bytes, dims 1..rank-1
  cuuint32_t box[3] = {key.b0, key.b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(key.ptr), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    delete m;
    set_error(RF_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d): ptr=%p dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u)",
              static_cast<int>(r), key.ptr, (unsigned long long)key.d0, (unsigned long long)key.d1,
              (unsigned long long)key.d2, (unsigned long long)key.s1, (unsigned long long)key.s2, key.b0, key.b1);
    return nullptr;
  }
  g_maps.emplace(key, m);
  return m;
}

const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                               uint32_t box_cols) {
  MapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.d0 = cols; k.d1 = rows; k.d2 = 1; k.s1 = ld_elems; k.s2 = 0; k.b0 = box_cols; k.b1 = box_rows;
  return lookup_or_encode(k, 2);
}

const CUtensorMap* get_tmap_3d(const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                               uint64_t batch_stride_elems, uint32_t box_rows) {
  MapKey k;
  std::memset(&k, 0, sizeof(k));
  k.ptr = ptr; k.d0 = cols; k.d1 = rows; k.d2 = batch; k.s1 = ld_elems; k.s2 = batch_stride_elems; k.b0 = 64;
  k.b1 = box_rows;
  return lookup_or_encode(k, 3);
}

}  // namespace rf

extern "C" {
const char* rf_last_error(void) { return rf::g_err; }
int rf_version(void) { return 102; }   // 102: rf_attn_args.keepbits, rf_gemm_args.drop_mask, rf_layernorm_bwd(drop_mask)
unsigned long long rf_launch_count(void) { return rf::g_launches.load(); }
int rf_set_dropout_nonce(const unsigned long long* nonce_dev, rf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  RF_REQUIRE(nonce_dev != nullptr, "rf_set_dropout_nonce: null pointer");
  for (int i = 0; i < rf::g_n_nonce_loaders; ++i) {
    rf::g_nonce_loaders[i](nonce_dev, stream);
    const int rc = rf::check_launch("rf_set_dropout_nonce");
    if (rc) return rc;
  }
  return RF_OK;
}
}

namespace rf { void set_row_activity(const RowActivity& a); }

extern "C" int rf_set_row_activity(const uint8_t* tile_flags, long long rows, const int32_t* qtile_list,
                                   const int32_t* n_qtiles) {
  if (tile_flags == nullptr) {
    rf::set_row_activity(rf::RowActivity{nullptr, 0, nullptr, nullptr});
    return RF_OK;
  }
  RF_REQUIRE(rows > 0 && rows % 256 == 0 && qtile_list && n_qtiles, "rf_set_row_activity: rows must be a positive multiple of 256");
  rf::set_row_activity(rf::RowActivity{tile_flags, rows, qtile_list, n_qtiles});
  return RF_OK;
}
