// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) as a function of N and of where the A
// operand lives (shared memory "SS" vs tensor memory "TS"), one CTA per SM, all SMs busy.  Development aid for the
// attention kernels: small-N SS MMAs re-read the 4 KB A slice for every instruction and are shared-memory bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I apla_b200/csrc tools/mma_microbench.cu -o /tmp/mb
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

// mode 0: SS, A K-major / B K-major   1: TS, B K-major   2: SS, A K-major / B MN-major   3: TS, B MN-major
// 4: SS, A MN-major (M=128 = two boxes 16 KB apart) / B MN-major
__global__ void __launch_bounds__(128, 1) bench(int mode, int N, int iters, int nacc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
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
    const bool b_mn = mode >= 2;
    const uint32_t idesc = make_idesc_bf16(128, N, mode == 4 ? 1 : 0, b_mn ? 1 : 0);
    long long t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t bd = b_mn ? DESC_LO_MN + b_lo + kk * 128 : DESC_LO_K + b_lo + 2 * kk;
          // nacc independent accumulators used round-robin: separates per-instruction cost from the latency of a
          // dependent accumulation chain
          const uint32_t d = tmem + (nacc == 1 ? 0 : ((it * 4 + kk) & (nacc - 1)) * 64);
          if (mode == 0 || mode == 2) umma_ss(d, DESC_LO_K + a_lo + 2 * kk, bd, idesc, 1);
          else if (mode == 4) umma_ss(d, DESC_LO_MN + a_lo + kk * 128, bd, idesc, 1);
          else umma_ts(d, tmem + 256 + kk * 8, bd, idesc, 1);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

// nw issuer warps (one per SM sub-partition), each with its own accumulator and commit barrier: is the per-instruction
// cost a property of the issuing thread or of the tensor pipe?
__global__ void __launch_bounds__(128, 1) bench_multi(int mode, int N, int iters, int nw, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_barrier_init();
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
  const long long t0 = clock64();
  if (warp < nw) {
    const bool leader = elect_one();
    const uint32_t a_lo = (smem_u32(smem) >> 4) + warp * 1024, b_lo = (smem_u32(smem + 64 * 1024) >> 4) + warp * 1024;
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    const uint32_t d = tmem + warp * 64;
    if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (mode == 0) umma_ss(d, DESC_LO_K + a_lo + 2 * kk, DESC_LO_K + b_lo + 2 * kk, idesc, 1);
          else umma_ts(d, tmem + 256 + warp * 32 + kk * 8, DESC_LO_K + b_lo + 2 * kk, idesc, 1);
        }
      }
      umma_commit(&bar[warp]);
    }
    __syncwarp();
    mbar_wait(&bar[warp], 0);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[] = {"SS  A=K-major B=K-major ", "TS  B=K-major           ", "SS  A=K-major B=MN-major", "TS  B=MN-major          ",
                         "SS  A=MN-major B=MN-major"};
  const int iters = 2000;
  for (int mode = 0; mode < 5; ++mode) {
    for (int N : {16, 64, 128, 256}) for (int nacc : {1, 2, 4}) {
      if (mode >= 2 && N > 64) continue;   // MN-major B wider than one 64-element box is not laid out in this test
      if (nacc > 1 && N > 64) continue;
      bench<<<148, 128, smem>>>(mode, N, iters, nacc, d);
      long long c = 0;
      cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        printf("%s N=%3d: %s\n", names[mode], N, cudaGetErrorString(e));
        return 1;
      }
      const double per = double(c) / (iters * 4);
      printf("%s N=%3d acc=%d: %7.1f cycles per MMA (floor %5.1f)  %6.1f B/clk smem\n", names[mode], N, nacc, per, 128.0 * N / 256,
             ((mode == 1 || mode == 3 ? 0 : 4096) + 32.0 * N) / per);
    }
  }
  cudaFuncSetAttribute(bench_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {16, 64, 128})
      for (int nw : {1, 2, 4}) {
        bench_multi<<<148, 128, smem>>>(mode, N, iters, nw, d);
        long long c = 0;
        cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
          printf("multi: %s\n", cudaGetErrorString(e));
          return 1;
        }
        printf("%s N=%3d, %d issuer warps: %7.1f cycles per MMA (aggregate)\n", mode ? "TS" : "SS", N, nw,
               double(c) / (iters * 4 * nw));
      }
  return 0;
}
