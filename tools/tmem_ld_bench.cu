// tcgen05.ld throughput: W warps (one per TMEM lane quadrant, W = 4 or 8) each load 32 lanes x 32 columns (4 KB) per
// instruction, `depth` loads in flight before each wait.  Prints bytes per SM clock.
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace apla;
__global__ void __launch_bounds__(256, 1) bench(int nwarps, int depth, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc<1>(&slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_addr = tmem + (uint32_t((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
      uint32_t v0[32], v1[32], v2[32];
      tmem_ld_32x32(lane_addr, v0);
      if (depth > 1) tmem_ld_32x32(lane_addr + 32, v1);
      if (depth > 2) tmem_ld_32x32(lane_addr + 64, v2);
      tmem_ld_wait();
      acc += __uint_as_float(v0[0]) + __uint_as_float(v0[31]);
      if (depth > 1) acc += __uint_as_float(v1[0]) + __uint_as_float(v1[31]);
      if (depth > 2) acc += __uint_as_float(v2[0]) + __uint_as_float(v2[31]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 123.456f) sink[threadIdx.x] = acc;
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
  cudaMalloc(&d, 8);
  cudaMalloc(&sink, 4096);
  const int iters = 2000;
  for (int nw : {1, 4, 8})
    for (int depth : {1, 3}) {
      bench<<<148, 256>>>(nw, depth, iters, d, sink);
      long long c = 0;
      if (cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
      printf("%d warps, %d loads in flight: %6.1f cycles per 4 KB load per warp, %7.1f B/clk/SM\n", nw, depth,
             double(c) / (iters * depth), 4096.0 * nw * depth * iters / double(c));
    }
  return 0;
}
