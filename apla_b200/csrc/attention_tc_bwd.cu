// tcgen05 / TMEM attention backward for head_dim 64: two kernels, no atomics, deterministic.
//
//   attn_bwd_dq_tc    one CTA per 128-row q tile, loops over 64-row kv tiles:
//                       S = Q K^T, dP = dO V^T (smem x smem -> TMEM);  dS = P o (dP - delta) (registers, P from LSE);
//                       dQ += dS K  (A = dS read from TMEM, B = the K tile re-read MN-major)
//   attn_bwd_dkdv_tc  one CTA per 128-row kv tile, loops over 64-row q tiles:
//                       S^T = K Q^T, dP^T = V dO^T;  P^T, dS^T written back to TMEM in place;
//                       dV += P^T dO,  dK += dS^T Q  (A from TMEM, B = the dO / Q tiles re-read MN-major)
// Both recompute the scores (7 MMA units instead of the algorithmic 5) so that dQ needs no cross-CTA reduction.
// One thread per TMEM lane owns a row of the score tile; the per-column LSE / delta of the kv-stationary kernel are
// staged in shared memory by the producer warp.  TMEM: 256 columns per CTA, two CTAs per SM.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace atb {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int kThreads = 192;
constexpr uint32_t TILE128 = 128 * 128, TILE64 = 64 * 128;

__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct SeqInfo {
  int row_start, n;
};
__device__ __forceinline__ SeqInfo seq_info(const int* cu, int b, int n_fixed) {
  SeqInfo s;
  if (cu) {
    s.row_start = cu[b];
    s.n = cu[b + 1] - s.row_start;
  } else {
    s.row_start = b * n_fixed;
    s.n = n_fixed;
  }
  return s;
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int k) { return make_sdesc_sw128(base + k * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int k) { return make_sdesc_sw128(base + k * 2048, 16384, 1024); }

// 64 fp32 accumulator columns of this thread's row -> bf16 -> 128 contiguous bytes in global memory
__device__ __forceinline__ void store_row64(uint32_t taddr, __nv_bfloat16* dst, float mul, bool store) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t ov[32];
    tmem_ld_32x32(taddr + c * 32, ov);
    tmem_ld_wait();
    if (store) {
      uint4* d4 = reinterpret_cast<uint4*>(dst) + c * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(ov[8 * j + i]) * mul;
        d4[j] = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dQ:  TMEM  S [0,64)   dP [64,128) (dS bf16 pairs in place, [64,96))   dQ [128,192)
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t DQ_SMEM = 1024 + 2 * TILE128 + 4 * TILE64 + 128;
constexpr uint32_t DQ_COL_DP = 64, DQ_COL_DQ = 128, DQ_TMEM_COLS = 256;

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tma_q128, const __grid_constant__ CUtensorMap tma_kv64,
                      const __grid_constant__ CUtensorMap tma_do128, const float* __restrict__ lse,
                      const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                      const int* __restrict__ cu_seqlens, int n_fixed, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + TILE128;
  uint8_t* sK = smem + 2 * TILE128;               // [2]
  uint8_t* sV = smem + 2 * TILE128 + 2 * TILE64;  // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TILE128 + 4 * TILE64);
  uint64_t* bar_q = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sp_full = bars + 5;
  uint64_t* ds_full = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const SeqInfo sq = seq_info(cu_seqlens, b, n_fixed);
  const int q0 = blockIdx.x * 128;
  if (q0 >= sq.n) return;
  const int D = H * 64;
  const int nkv = (sq.n + 63) / 64;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tma_q128);
    tma_prefetch_desc(&tma_kv64);
    tma_prefetch_desc(&tma_do128);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(ds_full, 4);
    mbar_init(dq_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc<1>(tmem_slot, DQ_TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, 2 * TILE128);
      tma_load_2d(sQ, &tma_q128, bar_q, h * 64, sq.row_start + q0);
      tma_load_2d(sdO, &tma_do128, bar_q, h * 64, sq.row_start + q0);
      for (int t = 0; t < nkv; ++t) {
        const int s = t & 1;
        mbar_wait(&kv_empty[s], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE64);
        tma_load_2d(sK + s * TILE64, &tma_kv64, &kv_full[s], D + h * 64, sq.row_start + t * 64);
        tma_load_2d(sV + s * TILE64, &tma_kv64, &kv_full[s], 2 * D + h * 64, sq.row_start + t * 64);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(bar_q, 0);
      const uint32_t q_base = smem_u32(sQ), do_base = smem_u32(sdO);
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
      for (int t = 0; t < nkv; ++t) {
        const int s = t & 1;
        const int valid = min(64, sq.n - t * 64);
        const int n_mma = (valid + 15) & ~15;
        mbar_wait(&kv_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sK + s * TILE64), v_base = smem_u32(sV + s * TILE64);
        const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem, desc_kmajor(q_base, k), desc_kmajor(k_base, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16<1>(tmem + DQ_COL_DP, desc_kmajor(do_base, k), desc_kmajor(v_base, k), idesc_s, k > 0);
        umma_commit(sp_full);
        mbar_wait(ds_full, t & 1);
        tc_fence_after();
        for (int k = 0; k < n_mma / 16; ++k)
          umma_ts_bf16(tmem + DQ_COL_DQ, tmem + DQ_COL_DP + k * 8, desc_mnmajor(k_base, k), idesc_acc,
                       (t > 0 || k > 0) ? 1u : 0u);
        umma_commit(&kv_empty[s]);
      }
      umma_commit(dq_full);
    }
  } else {
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    const float sl2 = scale * LOG2E;
    const bool row_ok = q0 + row < sq.n;
    const size_t tok = size_t(sq.row_start + q0 + row);
    const float lse2 = row_ok ? lse[tok * H + h] * LOG2E : 0.f;
    const float dl = row_ok ? delta[tok * H + h] : 0.f;
    for (int t = 0; t < nkv; ++t) {
      const int valid = min(64, sq.n - t * 64);
      const int n_mma = (valid + 15) & ~15;
      mbar_wait(sp_full, t & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < n_mma) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(lane_addr + c * 32, sv);
          tmem_ld_32x32(lane_addr + DQ_COL_DP + c * 32, dv);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c * 32 + 2 * j;
            const float p0 = exp2f(__uint_as_float(sv[2 * j]) * sl2 - lse2);
            const float p1 = exp2f(__uint_as_float(sv[2 * j + 1]) * sl2 - lse2);
            const float d0 = col < valid ? p0 * (__uint_as_float(dv[2 * j]) - dl) : 0.f;
            const float d1 = col + 1 < valid ? p1 * (__uint_as_float(dv[2 * j + 1]) - dl) : 0.f;
            pk[j] = pack_bf16(d0, d1);
          }
          tmem_st_32x16(lane_addr + DQ_COL_DP + c * 16, pk);   // chunk c of dS overwrites dP columns already consumed
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_full, 0);
    tc_fence_after();
    store_row64(lane_addr + DQ_COL_DQ, dqkv + tok * (3 * D) + h * 64, scale, row_ok);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, DQ_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dK, dV:  TMEM  S^T [0,64) (P^T pairs in place [0,32))   dP^T [64,128) (dS^T pairs in place [64,96))
//                dV [128,192)   dK [192,256)
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t KV_SMEM = 1024 + 2 * TILE128 + 4 * TILE64 + 2 * 2 * 64 * 4 + 128;
constexpr uint32_t KV_COL_DP = 64, KV_COL_DV = 128, KV_COL_DK = 192, KV_TMEM_COLS = 256;

__global__ void __launch_bounds__(kThreads, 2)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tma_kv128, const __grid_constant__ CUtensorMap tma_q64,
                        const __grid_constant__ CUtensorMap tma_do64, const float* __restrict__ lse,
                        const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                        const int* __restrict__ cu_seqlens, int n_fixed, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE128;
  uint8_t* sQ = smem + 2 * TILE128;                // [2] 64-row tiles
  uint8_t* sdO = smem + 2 * TILE128 + 2 * TILE64;  // [2]
  float* sL = reinterpret_cast<float*>(smem + 2 * TILE128 + 4 * TILE64);  // [2][64] lse * log2e (+inf past the end)
  float* sD = sL + 2 * 64;                                                // [2][64] delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 2 * 64);
  uint64_t* bar_kv = bars;
  uint64_t* q_full = bars + 1;   // [2] TMA bytes of the Q / dO tiles
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* ld_full = bars + 5;  // [2] lse / delta staged
  uint64_t* sp_full = bars + 7;
  uint64_t* pd_full = bars + 8;
  uint64_t* acc_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const SeqInfo sq = seq_info(cu_seqlens, b, n_fixed);
  const int kv0 = blockIdx.x * 128;
  if (kv0 >= sq.n) return;
  const int D = H * 64;
  const int nq = (sq.n + 63) / 64;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tma_kv128);
    tma_prefetch_desc(&tma_q64);
    tma_prefetch_desc(&tma_do64);
    mbar_init(bar_kv, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&ld_full[i], 1);
    }
    mbar_init(sp_full, 1);
    mbar_init(pd_full, 4);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc<1>(tmem_slot, KV_TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // whole warp: lane 0 drives TMA, all lanes stage the per-row statistics of each q tile
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_kv, 2 * TILE128);
      tma_load_2d(sK, &tma_kv128, bar_kv, D + h * 64, sq.row_start + kv0);
      tma_load_2d(sV, &tma_kv128, bar_kv, 2 * D + h * 64, sq.row_start + kv0);
    }
    for (int t = 0; t < nq; ++t) {
      const int s = t & 1;
      if (lane == 0) mbar_wait(&q_empty[s], ((t >> 1) & 1) ^ 1);
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(&q_full[s], 2 * TILE64);
        tma_load_2d(sQ + s * TILE64, &tma_q64, &q_full[s], h * 64, sq.row_start + t * 64);
        tma_load_2d(sdO + s * TILE64, &tma_do64, &q_full[s], h * 64, sq.row_start + t * 64);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = t * 64 + lane + 32 * i;
        const bool ok = r < sq.n;
        const size_t idx = size_t(sq.row_start + (ok ? r : 0)) * H + h;
        sL[s * 64 + lane + 32 * i] = ok ? __ldg(lse + idx) * LOG2E : INFINITY;
        sD[s * 64 + lane + 32 * i] = ok ? __ldg(delta + idx) : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ld_full[s]);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(bar_kv, 0);
      const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV);
      const uint32_t idesc_acc = make_idesc_bf16(128, 64, 0, 1);
      for (int t = 0; t < nq; ++t) {
        const int s = t & 1;
        const int valid = min(64, sq.n - t * 64);
        const int n_mma = (valid + 15) & ~15;
        mbar_wait(&q_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t q_base = smem_u32(sQ + s * TILE64), do_base = smem_u32(sdO + s * TILE64);
        const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem, desc_kmajor(k_base, k), desc_kmajor(q_base, k), idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16<1>(tmem + KV_COL_DP, desc_kmajor(v_base, k), desc_kmajor(do_base, k), idesc_s, k > 0);
        umma_commit(sp_full);
        mbar_wait(pd_full, t & 1);
        tc_fence_after();
        for (int k = 0; k < n_mma / 16; ++k) {
          umma_ts_bf16(tmem + KV_COL_DV, tmem + k * 8, desc_mnmajor(do_base, k), idesc_acc, (t > 0 || k > 0) ? 1u : 0u);
          umma_ts_bf16(tmem + KV_COL_DK, tmem + KV_COL_DP + k * 8, desc_mnmajor(q_base, k), idesc_acc,
                       (t > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&q_empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int row = warp * 32 + lane;  // kv row of this tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    const float sl2 = scale * LOG2E;
    for (int t = 0; t < nq; ++t) {
      const int s = t & 1;
      const int valid = min(64, sq.n - t * 64);
      const int n_mma = (valid + 15) & ~15;
      mbar_wait(&ld_full[s], (t >> 1) & 1);
      mbar_wait(sp_full, t & 1);
      tc_fence_after();
      // (explicit ld.shared: through the generic pointers these 32 statistic loads per chunk and thread went to the
      //  local/global queue instead of the shared-memory pipe -- see ptx.cuh)
      const uint32_t L = smem_u32(sL) + s * 256, Dl = smem_u32(sD) + s * 256;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < n_mma) {
          uint32_t sv[32], dv[32];
          tmem_ld_32x32(lane_addr + c * 32, sv);
          tmem_ld_32x32(lane_addr + KV_COL_DP + c * 32, dv);
          tmem_ld_wait();
          uint32_t pp[16], pd[16];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 l4 = lds_f4(L + (c * 32 + 4 * j4) * 4);
            const float4 d4 = lds_f4(Dl + (c * 32 + 4 * j4) * 4);
            const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const int j = 2 * j4 + h2;
              const int col = c * 32 + 2 * j;
              const float p0 = col < valid ? exp2f(__uint_as_float(sv[2 * j]) * sl2 - lv[2 * h2]) : 0.f;
              const float p1 = col + 1 < valid ? exp2f(__uint_as_float(sv[2 * j + 1]) * sl2 - lv[2 * h2 + 1]) : 0.f;
              const float e0 = col < valid ? p0 * (__uint_as_float(dv[2 * j]) - dl4[2 * h2]) : 0.f;
              const float e1 = col + 1 < valid ? p1 * (__uint_as_float(dv[2 * j + 1]) - dl4[2 * h2 + 1]) : 0.f;
              pp[j] = pack_bf16(p0, p1);
              pd[j] = pack_bf16(e0, e1);
            }
          }
          tmem_st_32x16(lane_addr + c * 16, pp);
          tmem_st_32x16(lane_addr + KV_COL_DP + c * 16, pd);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pd_full);
    }
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool row_ok = kv0 + row < sq.n;
    __nv_bfloat16* base = dqkv + size_t(sq.row_start + kv0 + row) * (3 * D) + h * 64;
    store_row64(lane_addr + KV_COL_DK, base + D, scale, row_ok);
    store_row64(lane_addr + KV_COL_DV, base + 2 * D, 1.0f, row_ok);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, KV_TMEM_COLS);
  }
}

}  // namespace atb

int attn_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* delta, void* dqkv,
                const int* cu_seqlens, int num_seqs, int max_seqlen, int total_tokens, int H, float scale,
                cudaStream_t stream) {
  using namespace atb;
  const int D = H * 64;
  CUtensorMap q128, q64, do128, do64;
  if (int rc = make_tmap_2d(&q128, qkv, 2, total_tokens, 3 * D, 3 * D, 128, 64, true)) return rc;
  if (int rc = make_tmap_2d(&q64, qkv, 2, total_tokens, 3 * D, 3 * D, 64, 64, true)) return rc;
  if (int rc = make_tmap_2d(&do128, dout, 2, total_tokens, D, D, 128, 64, true)) return rc;
  if (int rc = make_tmap_2d(&do64, dout, 2, total_tokens, D, D, 64, 64, true)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DQ_SMEM));
    APLA_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KV_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(max_seqlen, 128), num_seqs * H);
  attn_bwd_dkdv_tc_kernel<<<grid, kThreads, KV_SMEM, stream>>>(q128, q64, do64, lse, delta,
                                                               reinterpret_cast<__nv_bfloat16*>(dqkv), cu_seqlens,
                                                               max_seqlen, H, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  attn_bwd_dq_tc_kernel<<<grid, kThreads, DQ_SMEM, stream>>>(q128, q64, do128, lse, delta,
                                                             reinterpret_cast<__nv_bfloat16*>(dqkv), cu_seqlens,
                                                             max_seqlen, H, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
