// Microbenchmark 2: the forward-attention MMA mix (6 x TS N=64 accumulate + 4 x SS N=96) issued back to back by one
// thread, (a) alone, (b) while 4 / 8 other warps stream tcgen05.ld over 96 columns, (c) ... ld + st.  Is the
// in-kernel cost per MMA (~120 cycles) tensor-memory contention or a property of the instruction mix?
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace apla;
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t DESC_LO_K = (16u >> 4) << 16;
constexpr uint32_t DESC_LO_MN = (16384u >> 4) << 16;
__device__ __forceinline__ void umma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_t, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %4, 0;\n\tmov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d),
      "r"(a_t), "r"(b_lo), "r"(idesc), "r"(acc), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// pattern 0: 6 TS(N=64) + 4 SS(N=96);  1: 10 SS (N=64);  2: 10 TS (N=64);  3: 4 SS(N=96) only;  4: 6 TS only
// noise 0: none; 1: 4 warps ld; 2: 8 warps ld; 3: 8 warps ld+st; 4: 8 warps doing exp2 math only (MUFU + FMA)
__global__ void __launch_bounds__(320, 1) bench(int pattern, int noise, int iters, long long* out, float* sink, int commits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2[0], 1);
    mbar_init(&bar2[1], 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 0) {
    tmem_alloc<1>(&slot, 512);
    tmem_relinquish<1>();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t a_lo = smem_u32(smem) >> 4, b_lo = smem_u32(smem + 64 * 1024) >> 4;
    const uint32_t id96 = make_idesc_bf16(128, 96, 0, 0), id64 = make_idesc_bf16(128, 64, 0, 1), id64k = make_idesc_bf16(128, 64, 0, 0);
    long long t0 = clock64();
    int n_mma = 0;
    if (leader) {
      for (int it = 0; it < iters; ++it) {
        if (pattern == 0 || pattern == 4) {
#pragma unroll
          for (int kk = 0; kk < 6; ++kk) umma_ts(tmem + 192, tmem + kk * 8, DESC_LO_MN + b_lo + kk * 128, id64, 1);
          if (commits) umma_commit(&bar2[0]);
        }
        if (pattern == 0 || pattern == 3) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + 96, DESC_LO_K + a_lo + 2 * kk, DESC_LO_K + b_lo + 2 * kk, id96, kk > 0);
          if (commits) umma_commit(&bar2[1]);
        }
        if (pattern == 1) {
#pragma unroll
          for (int kk = 0; kk < 10; ++kk) umma_ss(tmem + 192, DESC_LO_K + a_lo + 2 * (kk & 3), DESC_LO_K + b_lo + 2 * (kk & 3), id64k, 1);
        }
        if (pattern == 2) {
#pragma unroll
          for (int kk = 0; kk < 10; ++kk) umma_ts(tmem + 192, tmem + (kk & 3) * 8, DESC_LO_MN + b_lo + (kk & 3) * 128, id64, 1);
        }
      }
      umma_commit(&bar);
    }
    n_mma = iters * (pattern == 0 ? 10 : pattern == 3 ? 4 : pattern == 4 ? 6 : 10);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) {
      out[0] = t1 - t0;
      out[1] = n_mma;
    }
    stop = 1;
  } else if (warp >= 2) {
    const int nw = noise == 1 ? 4 : (noise >= 2 ? 8 : 0);
    if (warp - 2 < nw) {
      const uint32_t lane_addr = tmem + (uint32_t((warp & 3) * 32) << 16) + 256 + ((warp - 2) >> 2) * 96;
      float acc = 0.f;
      while (!stop) {
        if (noise == 4) {
#pragma unroll
          for (int i = 0; i < 96; ++i) acc += exp2f(acc * 0.001f + i);
        } else {
#pragma unroll
          for (int pc = 0; pc < 3; ++pc) {
            uint32_t v[32];
            tmem_ld_32x32(lane_addr + pc * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i]);
            if (noise == 3) {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) pk[i] = v[2 * i] ^ v[2 * i + 1];
              tmem_st_32x16(lane_addr + pc * 16, pk);
            }
          }
          if (noise == 3) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
      }
      if (acc == 123.456f) sink[threadIdx.x] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}
int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 16);
  cudaMalloc(&sink, 4096);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* pn[] = {"6 TS(N=64) + 4 SS(N=96)", "10 SS N=64", "10 TS N=64", "4 SS N=96", "6 TS N=64"};
  const char* nn[] = {"alone", "4 warps tcgen05.ld", "8 warps tcgen05.ld", "8 warps ld+st", "8 warps exp2 math"};
  for (int p = 0; p < 1; ++p)
    for (int n = 0; n < 2; ++n) {
      const int commits = n;
      bench<<<148, 320, smem>>>(p, 0, 1000, d, sink, commits);
      long long c[2] = {0, 0};
      cudaError_t e = cudaMemcpy(c, d, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        printf("%s\n", cudaGetErrorString(e));
        return 1;
      }
      printf("%-26s commits=%d: %7.1f cycles per MMA\n", pn[p], commits, double(c[0]) / double(c[1]));
    }
  return 0;
}
