// Host-side helpers shared by the translation units of librecformer_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/recformer_b200.h"

namespace rf {

// Error channel of the C ABI: every entry point returns 0 or a negative code and leaves a
// human-readable message retrievable with rf_last_error() (thread local).
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError() -> rf error

#define RF_REQUIRE(cond, ...)                                  \
  do {                                                         \
    if (!(cond)) return ::rf::set_error(RF_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define RF_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ::rf::set_error(RF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__));  \
  } while (0)

// bf16 2-D row-major tensor map with 128B swizzle; box = (64 elements, box_rows).  Cached per
// (ptr, shape, box).  Returns nullptr on failure (error already set).
const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                               uint32_t box_cols = 64);
// bf16 3-D (batch, rows, cols) map: box = (64 cols, box_rows rows, 1 batch); out-of-range rows
// (negative or >= rows) are zero-filled, which the band kernels rely on at sequence edges.
const CUtensorMap* get_tmap_3d(const void* ptr, uint64_t batch, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                               uint64_t batch_stride_elems, uint32_t box_rows);

// Launch context "row activity" (rf_set_row_activity): when set, the token-major kernels of the encoder skip the
// 256-row tiles whose flag is 0 (tiles made of padding only).  Thread-local, consulted at launch time.
struct RowActivity {
  const uint8_t* flags;        // [rows / 256], device; nullptr = everything active
  long long rows;              // B * L of the activation matrices the flags describe
  const int32_t* qtiles;       // compact list of the active 128-row query tiles (b * (L / 128) + tile), device
  const int32_t* n_qtiles;     // its length, device
};
const RowActivity& row_activity();

int sm_count();
// true exactly once per (flag, current CUDA device): guards the per-device cudaFuncSetAttribute calls
bool first_use_on_device(std::atomic<unsigned long long>* seen);

// Translation units whose kernels draw dropout masks register a loader for their copy of
// rf_dropout_nonce (rf_ptx.cuh); rf_set_dropout_nonce() runs every registered loader.
typedef void (*nonce_loader_fn)(const unsigned long long* dev_src, cudaStream_t stream);
int register_nonce_loader(nonce_loader_fn fn);

}  // namespace rf

#define RF_DEFINE_NONCE_LOADER(tag)                                                                       \
  namespace rf {                                                                                          \
  static __global__ void nonce_load_kernel_##tag(const unsigned long long* src) { rf_dropout_nonce = *src; } \
  static void nonce_loader_##tag(const unsigned long long* src, cudaStream_t stream) {                    \
    nonce_load_kernel_##tag<<<1, 1, 0, stream>>>(src);                                                    \
  }                                                                                                       \
  static const int nonce_registered_##tag = register_nonce_loader(nonce_loader_##tag);                    \
  }
