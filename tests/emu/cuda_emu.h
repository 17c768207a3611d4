// A minimal SIMT emulator for the CPU: enough of the CUDA execution model to run the row kernels of
// apla_b200/csrc/ssl.cu UNCHANGED (tests/emu/build_emu.py rewrites only the <<<...>>> launch syntax, the
// `extern __shared__` declaration and the two project includes).  TEST INFRASTRUCTURE ONLY.
//
// One block at a time; every CUDA thread of the block is an OS thread.  __syncthreads() is a std::barrier over the block,
// __shfl_xor_sync() an exchange through a per-warp buffer guarded by a per-warp barrier, `__shared__` variables are
// function-local statics (one instance, because blocks run one after another), dynamic shared memory is one buffer per
// launch.  Threads that return early drop out of their barriers, as exited CUDA threads do.  What this checks: indexing,
// reductions, barrier placement (a misplaced barrier deadlocks or trips the race the real GPU would have), the
// arithmetic in fp32 with libm's expf / logf.  What it cannot check: hardware scheduling, memory coalescing, fast-math
// intrinsics' rounding, performance.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static float* emu_dyn_smem = nullptr;

// ---- bf16 ------------------------------------------------------------------------------------------------------------
struct __nv_bfloat16 { uint16_t bits; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{uint16_t((u >> 16) | 0x40)};   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                                          // round to nearest even
  return __nv_bfloat16{uint16_t(u >> 16)};
}
static inline float __bfloat162float(__nv_bfloat16 h) {
  uint32_t u = uint32_t(h.bits) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
  return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)};
}

// ---- intrinsics ------------------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
#define __expf(x) expf(x)   // glibc declares the double-underscore names itself
#define __logf(x) logf(x)
using std::max;
using std::min;

// ---- block / warp synchronisation ------------------------------------------------------------------------------------
struct EmuBlock {
  std::unique_ptr<std::barrier<>> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<float> warp_buf;   // [warps][32]
};
static EmuBlock* emu_block = nullptr;
static inline unsigned emu_tid() { return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z); }

static inline void __syncthreads() { emu_block->block_bar->arrive_and_wait(); }
static inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  const unsigned t = emu_tid(), w = t >> 5, l = t & 31;
  float* buf = emu_block->warp_buf.data() + w * 32;
  buf[l] = v;
  emu_block->warp_bar[w]->arrive_and_wait();
  const float r = buf[l ^ unsigned(lane_mask)];
  emu_block->warp_bar[w]->arrive_and_wait();
  return r;
}

template <class F>
static void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  gridDim = grid;
  blockDim = block;
  const unsigned n = block.x * block.y * block.z;
  std::vector<float> dyn((smem_bytes + 3) / 4 + 1);
  emu_dyn_smem = dyn.data();
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        EmuBlock blk;
        blk.block_bar = std::make_unique<std::barrier<>>(n);
        const unsigned warps = (n + 31) / 32;
        for (unsigned w = 0; w < warps; ++w)
          blk.warp_bar.push_back(std::make_unique<std::barrier<>>(std::min(32u, n - w * 32)));
        blk.warp_buf.assign(size_t(warps) * 32, 0.f);
        emu_block = &blk;
        std::vector<std::thread> ts;
        ts.reserve(n);
        for (unsigned t = 0; t < n; ++t)
          ts.emplace_back([&, t] {
            threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            blockIdx = dim3(bx, by, bz);
            body();
            blk.warp_bar[t >> 5]->arrive_and_drop();   // an exited thread no longer takes part in barriers
            blk.block_bar->arrive_and_drop();
          });
        for (auto& th : ts) th.join();
      }
  emu_block = nullptr;
  emu_dyn_smem = nullptr;
}

// ---- the slice of the CUDA runtime / project headers that ssl.cu's launchers use ----------------------------------
typedef void* cudaStream_t;
typedef int cudaError_t;
static const cudaError_t cudaSuccess = 0;
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

namespace apla {
static char emu_error[512];
static inline void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(emu_error, sizeof emu_error, fmt, ap);
  va_end(ap);
}
static long long emu_launches = 0;
static inline void count_launch(int n = 1) { emu_launches += n; }
static inline int sm_count() { return 2; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
}  // namespace apla

#define APLA_CHECK(cond, ...)         \
  do {                                \
    if (!(cond)) {                    \
      ::apla::set_error(__VA_ARGS__); \
      return 1;                       \
    }                                 \
  } while (0)
#define APLA_CUDA(expr)                  \
  do {                                   \
    if ((expr) != cudaSuccess) return 2; \
  } while (0)
