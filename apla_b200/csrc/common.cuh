// Shared host-side helpers: error reporting for the C ABI, tensor-map encoding without linking libcuda.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace apla {

// last error message of the calling thread, exposed by apla_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define APLA_CHECK(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::apla::set_error(__VA_ARGS__); \
      return 1;                      \
    }                                \
  } while (0)

#define APLA_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::apla::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

// 2-D bf16/fp32 row-major tensor map: `rows` x `cols` (cols contiguous), leading dimension `ld` elements,
// box = box_rows x box_cols, 128-byte swizzle (box_cols * elt == 128 bytes) or none.
int make_tmap_2d(CUtensorMap* out, const void* base, int elt_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols, bool swizzle128);
// same with an explicit swizzle span in bytes (0, 64 or 128 = inner box bytes)
int make_tmap_2d_sw(CUtensorMap* out, const void* base, int elt_bytes, uint64_t rows, uint64_t cols, uint64_t ld,
                    uint32_t box_rows, uint32_t box_cols, int swizzle_bytes);

int sm_count();

// number of kernels launched by this library since load (bench.py reports the per-step delta as gpu_launches)
void count_launch(int n = 1);
long long launch_count();

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace apla
