// Data-parallel gradient all-reduce over NVLink peer memory, as ONE kernel that can be captured in the step's CUDA graph.
//
// Replaces the DDP reducer of the reference (src/defaults/wrappers.py:182-183: bucketed NCCL all-reduce of the trainable
// gradients) for the step engine's single fp32 gradient arena.  Every rank's arena lives in symmetric memory (the same
// allocation mapped into every process of the node; torch.distributed._symmetric_memory does the rendezvous), so a kernel
// can read and write its peers' arenas directly through NVSwitch:
//
//   ready barrier   every CTA tells the same-numbered CTA of every peer "my gradients are complete" (the kernel is
//                   stream-ordered behind the backward that produced them) and waits for the same from all of them
//   reduce-scatter  rank r owns the r-th 1/W of the slice: it reads that range from all W arenas (peer loads, 128-bit)
//                   and adds them in rank order 0..W-1 -- one owner per element, a fixed order: the result is
//                   bit-identical on every rank and from run to run
//   all-gather      ... and writes the sum into all W arenas (peer stores)
//   done barrier    nobody leaves before every peer's stores into its arena have been issued and fenced
//
// With NVLS (a multicast mapping of the arena) the switch does the adding: multimem.ld_reduce / multimem.st.
// Two-shot, in place, sum (the 1/world of DDP's mean is folded into the fused clip + AdamW kernel).  A 2 MB arena (C2)
// is latency-bound (~10 us at 8 GPUs); 30 MB (C3) moves 7/8 of the slice over NVLink in each direction.
// Flags: one symmetric uint32 array per communicator, [channel][phase][source rank][CTA]; values are launch epochs that
// only grow, so nothing is ever reset and a replayed CUDA graph keeps counting (the epoch of a (channel, CTA) lives in
// local device memory and is advanced by that CTA alone).  Waits are bounded: a rank that never shows up turns into a
// launch error (trap) on its peers, not a hung GPU.
#include "common.cuh"
#include "kernels.cuh"

namespace apla {
namespace dpar {

constexpr int kThreads = 128;          // small CTAs (128 threads x <= 88 registers = 11 K registers): they fit beside a
                                       // resident persistent-GEMM CTA (54 K registers), so the early slice really overlaps
constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 148;
constexpr int kChannels = 4;
constexpr long long kSpinLimit = 400000000LL;   // ~ a few seconds of polling

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flag slot of (channel, phase, source rank, cta) in a rank's flag array
__device__ __forceinline__ size_t flag_index(int channel, int phase, int src, int cta) {
  return ((size_t(channel) * 2 + phase) * kMaxWorld + src) * kMaxCtas + cta;
}

// every thread t < world: publish `epoch` in peer t's flags, then wait for peer t's flag in my array
__device__ __forceinline__ void cross_barrier(uint32_t* const* flags, int rank, int world, int channel, int phase,
                                              uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(flags[peer] + flag_index(channel, phase, rank, blockIdx.x), epoch);
    const uint32_t* mine = flags[rank] + flag_index(channel, phase, peer, blockIdx.x);
    long long spins = 0;
    // (epochs only grow; a wrapped comparison keeps working after 2^31 launches)
    while (int32_t(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > kSpinLimit) __trap();
    }
  }
  __syncthreads();
}

struct Peers {
  float* buf[kMaxWorld];
  uint32_t* flags[kMaxWorld];
};

__device__ __forceinline__ float4 multimem_ld_reduce(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// U elements (float4) per thread and trip, every load issued before the first add: the exchange is latency-bound (a peer
// load takes ~2 us), so bytes in flight are what set the bandwidth.  (First version: one element per trip, 60 GB/s with 16
// CTAs at 2 GPUs against NCCL's 400.)
template <int U, int PG>
__device__ __forceinline__ void reduce_range_peers(const Peers& peers, int world, int64_t offset, int64_t lo, int64_t hi) {
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  for (int64_t i0 = lo + int64_t(blockIdx.x) * kThreads + threadIdx.x; i0 < hi; i0 += stride * U) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p0 = 0; p0 < kMaxWorld; p0 += PG) {
      if (p0 < world) {
        float4 v[U][PG];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int p = 0; p < PG; ++p)
            v[u][p] = (p0 + p < world && i0 + u * stride < hi)
                          ? ld_peer(reinterpret_cast<const float4*>(peers.buf[p0 + p] + offset) + i0 + u * stride)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int p = 0; p < PG; ++p) {        // rank order 0..W-1: the same sum on every rank, every run
            acc[u].x += v[u][p].x; acc[u].y += v[u][p].y; acc[u].z += v[u][p].z; acc[u].w += v[u][p].w;
          }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < hi) {
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p)
          if (p < world) st_peer(reinterpret_cast<float4*>(peers.buf[p] + offset) + i0 + u * stride, acc[u]);
      }
  }
}

// NVLS: the switch adds the W copies (multimem.ld_reduce) and broadcasts the sum (multimem.st) -- 1 / W of the loads and
// stores of the peer loop, one owner per element as before
__device__ __forceinline__ void reduce_range_multimem(float* mc, int64_t offset, int64_t lo, int64_t hi) {
  constexpr int U = 4;
  const int64_t stride = int64_t(gridDim.x) * kThreads;
  float4* base = reinterpret_cast<float4*>(mc + offset);
  for (int64_t i0 = lo + int64_t(blockIdx.x) * kThreads + threadIdx.x; i0 < hi; i0 += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < hi) v[u] = multimem_ld_reduce(base + i0 + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < hi) multimem_st(base + i0 + u * stride, v[u]);
  }
}

__global__ void __maxnreg__(88)
arena_allreduce_kernel(const __grid_constant__ Peers peers, float* __restrict__ multicast, uint32_t* __restrict__ epochs,
                       int rank, int world, int64_t offset, int64_t count4, int64_t offset_b, int64_t count4_b,
                       int channel) {
  __shared__ uint32_t s_epoch;
  __shared__ uint32_t* s_flags[kMaxWorld];       // (indexed by thread: shared memory instead of a local-memory copy)
  if (threadIdx.x == 0) s_epoch = epochs[channel * kMaxCtas + blockIdx.x] + 1u;
  if (threadIdx.x < kMaxWorld) s_flags[threadIdx.x] = peers.flags[threadIdx.x];
  __syncthreads();
  const uint32_t epoch = s_epoch;
  cross_barrier(s_flags, rank, world, channel, 0, epoch);

  // my shard of each slice (float4 units), split over the CTAs; a second (small) slice rides on the same pair of barriers
  for (int part = 0; part < 2; ++part) {
    const int64_t n4 = part == 0 ? count4 : count4_b, off = part == 0 ? offset : offset_b;
    if (n4 == 0) continue;
    const int64_t per_rank = (n4 + world - 1) / world;
    const int64_t lo = min(n4, int64_t(rank) * per_rank), hi = min(n4, lo + per_rank);
    if (multicast != nullptr) reduce_range_multimem(multicast, off, lo, hi);
    else if (world <= 2) reduce_range_peers<4, 2>(peers, world, off, lo, hi);
    else reduce_range_peers<2, 4>(peers, world, off, lo, hi);
  }
  cross_barrier(s_flags, rank, world, channel, 1, epoch);
  if (threadIdx.x == 0) epochs[channel * kMaxCtas + blockIdx.x] = epoch;
}

}  // namespace dpar

int grad_arena_allreduce(float* const* peer_bufs, uint32_t* const* peer_flags, float* multicast, uint32_t* epochs, int rank,
                         int world, int64_t offset_floats, int64_t count_floats, int64_t offset_b, int64_t count_b,
                         int channel, int ctas, cudaStream_t stream) {
  using namespace dpar;
  APLA_CHECK(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "grad_arena_allreduce: rank %d / world %d", rank, world);
  APLA_CHECK(channel >= 0 && channel < kChannels, "grad_arena_allreduce: channel %d (0..%d)", channel, kChannels - 1);
  APLA_CHECK(offset_floats % 4 == 0 && count_floats % 4 == 0 && count_floats >= 0 && offset_b % 4 == 0 && count_b % 4 == 0 &&
                 count_b >= 0,
             "grad_arena_allreduce: offsets / counts (%lld / %lld, %lld / %lld floats) must be multiples of 4",
             (long long)offset_floats, (long long)count_floats, (long long)offset_b, (long long)count_b);
  if ((count_floats == 0 && count_b == 0) || world == 1) return 0;
  if (ctas <= 0) ctas = 32;
  if (ctas > kMaxCtas) ctas = kMaxCtas;
  Peers p;
  for (int i = 0; i < kMaxWorld; ++i) {
    p.buf[i] = i < world ? peer_bufs[i] : nullptr;
    p.flags[i] = i < world ? peer_flags[i] : nullptr;
    APLA_CHECK(i >= world || (p.buf[i] != nullptr && p.flags[i] != nullptr), "grad_arena_allreduce: null peer pointer %d", i);
  }
  arena_allreduce_kernel<<<ctas, kThreads, 0, stream>>>(p, multicast, epochs, rank, world, offset_floats, count_floats / 4,
                                                        offset_b, count_b / 4, channel);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int grad_arena_allreduce_flag_words() { return dpar::kChannels * 2 * dpar::kMaxWorld * dpar::kMaxCtas; }
int grad_arena_allreduce_epoch_words() { return dpar::kChannels * dpar::kMaxCtas; }

}  // namespace apla
