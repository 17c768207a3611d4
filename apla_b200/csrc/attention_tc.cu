// tcgen05 / TMEM fused softmax attention for head_dim 64 (forward + the two backward kernels).
//
// Common skeleton (one CTA per 128-row tile of the "stationary" operand, two CTAs co-resident per SM so that one
// CTA's tensor-core work overlaps the other's exponentials):
//   warp 4   TMA producer: operand tiles straight out of the packed qkv / dO matrices (128B swizzle)
//   warp 5   MMA issuer (one thread): score-type MMAs (A, B from shared memory) into TMEM, then the accumulate-type
//            MMAs whose A operand (P, dS; bf16) is read FROM TMEM where the compute warps left it
//   warps 0-3  one thread per TMEM lane = per row of the score tile: tcgen05.ld -> exp2 / products in registers ->
//            tcgen05.st of the packed bf16 operand; no shuffles, no shared-memory round trip for P / dS
// Forward keeps the running output in registers (online softmax: each KV tile's P.V lands in a fresh TMEM tile and is
// folded in with the rescale), so TMEM accumulators never need an in-place correction pass.
// Row / column tails: the score MMA's N is sized to the valid columns (multiple of 16), masked columns get P = 0;
// rows past the sequence end compute on whatever the tile holds and are simply not stored.
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace apla {
namespace atc {

constexpr float LOG2E = 1.4426950408889634f;
constexpr int kThreads = 192;

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

// K-major operand tile (rows x 64 bf16, 128 B per row, 128B swizzle): advance 16 elements = 32 B per MMA k-step
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int k) { return make_sdesc_sw128(base + k * 32, 16, 1024); }
// the same tile read as an MN-major operand (rows are the contraction index): 16 rows = 2048 B per MMA k-step
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int k) { return make_sdesc_sw128(base + k * 2048, 16384, 1024); }

// ------------------------------------------------------------------------------------------------------------------
// forward:  O = softmax(scale * Q K^T) V,  lse = log-sum-exp per row
//
// 128 q rows x 64 kv columns per step.  TMEM (128 columns, so four CTAs share an SM and hide each other's latencies):
//   [0,64)  S fp32, overwritten in place by P (bf16 pairs, [0,32))      [64,128)  O fp32, accumulated by the MMAs
// The running maximum is "lazy": P is formed against m_used, which is only raised (and O / l rescaled in TMEM) when a
// row's maximum has grown by more than 2^8 -- P then stays below 256, exact in fp32/bf16 terms, and the rescale pass,
// which would otherwise sit on the critical path of every step, practically never runs after the first tile.
// ------------------------------------------------------------------------------------------------------------------
constexpr int F_BQ = 128, F_BKV = 64;
constexpr uint32_t F_QTILE = 128 * 128, F_KVTILE = 64 * 128;
constexpr uint32_t F_SMEM = 1024 + F_QTILE + 4 * F_KVTILE + 128;
constexpr uint32_t F_TMEM_COLS = 128, F_COL_O = 64;

__global__ void __launch_bounds__(kThreads, 3)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_kv,
                   __nv_bfloat16* __restrict__ out, float* __restrict__ lse, const int* __restrict__ cu_seqlens,
                   int n_fixed, int H, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + F_QTILE;                 // [2]
  uint8_t* sV = smem + F_QTILE + 2 * F_KVTILE;  // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_QTILE + 4 * F_KVTILE);
  uint64_t* bar_q = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / H, h = blockIdx.y % H;
  const SeqInfo sq = seq_info(cu_seqlens, b, n_fixed);
  const int q0 = blockIdx.x * F_BQ;
  if (q0 >= sq.n) return;
  const int D = H * 64;
  const int nkv = (sq.n + F_BKV - 1) / F_BKV;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_kv);
    mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) {
    tmem_alloc<1>(tmem_slot, F_TMEM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_q, F_QTILE);
      tma_load_2d(sQ, &tma_q, bar_q, h * 64, sq.row_start + q0);
      for (int t = 0; t < nkv; ++t) {
        const int s = t & 1;
        mbar_wait(&kv_empty[s], ((t >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * F_KVTILE);
        tma_load_2d(sK + s * F_KVTILE, &tma_kv, &kv_full[s], D + h * 64, sq.row_start + t * F_BKV);
        tma_load_2d(sV + s * F_KVTILE, &tma_kv, &kv_full[s], 2 * D + h * 64, sq.row_start + t * F_BKV);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      mbar_wait(bar_q, 0);
      const uint32_t q_base = smem_u32(sQ);
      const uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      for (int t = 0; t < nkv; ++t) {
        const int s = t & 1;
        const int valid = min(F_BKV, sq.n - t * F_BKV);
        const int n_mma = (valid + 15) & ~15;
        mbar_wait(&kv_full[s], (t >> 1) & 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sK + s * F_KVTILE), v_base = smem_u32(sV + s * F_KVTILE);
        const uint32_t idesc_s = make_idesc_bf16(128, n_mma, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem, desc_kmajor(q_base, k), desc_kmajor(k_base, k), idesc_s, k > 0);
        umma_commit(s_full);
        mbar_wait(p_full, t & 1);
        tc_fence_after();
        for (int k = 0; k < n_mma / 16; ++k)
          umma_ts_bf16(tmem + F_COL_O, tmem + k * 8, desc_mnmajor(v_base, k), idesc_o, (t > 0 || k > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    const int row = warp * 32 + lane;                   // row of the q tile == TMEM lane
    const uint32_t lane_addr = tmem + (uint32_t(warp * 32) << 16);
    const float sl2 = scale * LOG2E;
    float m_used = -INFINITY, l_run = 0.f;
    for (int t = 0; t < nkv; ++t) {
      const int valid = min(F_BKV, sq.n - t * F_BKV);
      const int n_mma = (valid + 15) & ~15;
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld_32x32(lane_addr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      if (n_mma > 32) tmem_ld_32x32(lane_addr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      float mx = -INFINITY;
      if (valid == F_BKV) {
#pragma unroll
        for (int j = 0; j < 64; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 64; ++j)
          if (j < valid) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      if (t == 0) {
        m_used = mx;
      } else {
        const bool need = (mx - m_used) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // rare: raise the reference maximum of the rows that need it and rescale their O / l in place
          const float f = need ? exp2f((m_used - mx) * sl2) : 1.0f;
          if (need) m_used = mx;
          l_run *= f;
          mbar_wait(o_full, (t - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[16];
            tmem_ld_32x16(lane_addr + F_COL_O + c * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) ov[j] = __float_as_uint(__uint_as_float(ov[j]) * f);
            tmem_st_32x16(lane_addr + F_COL_O + c * 16, ov);
          }
        }
      }
      const float ms = m_used * sl2;
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c * 32 < n_mma) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c * 32 + 2 * j;
            const float p0 = col < valid ? exp2f(__uint_as_float(v[col]) * sl2 - ms) : 0.f;
            const float p1 = col + 1 < valid ? exp2f(__uint_as_float(v[col + 1]) * sl2 - ms) : 0.f;
            rs += p0 + p1;
            pk[j] = pack_bf16(p0, p1);
          }
          tmem_st_32x16(lane_addr + c * 16, pk);
        }
      }
      l_run += rs;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, (nkv - 1) & 1);
    tc_fence_after();
    const bool store = q0 + row < sq.n;
    const float inv = 1.f / l_run;
    uint4* dst = reinterpret_cast<uint4*>(out + size_t(sq.row_start + q0 + row) * D + h * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld_32x32(lane_addr + F_COL_O + c * 32, ov);
      tmem_ld_wait();
      if (store) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(ov[8 * j + i]) * inv;
          dst[c * 4 + j] = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
        }
      }
    }
    if (store) lse[size_t(sq.row_start + q0 + row) * H + h] = m_used * scale + logf(l_run);
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, F_TMEM_COLS);
  }
}

}  // namespace atc

int attn_fwd_tc(const void* qkv, void* out, float* lse, const int* cu_seqlens, int num_seqs, int max_seqlen,
                int total_tokens, int H, float scale, cudaStream_t stream) {
  using namespace atc;
  CUtensorMap tq, tkv;
  if (int rc = make_tmap_2d(&tq, qkv, 2, total_tokens, 3 * H * 64, 3 * H * 64, 128, 64, true)) return rc;
  if (int rc = make_tmap_2d(&tkv, qkv, 2, total_tokens, 3 * H * 64, 3 * H * 64, 64, 64, true)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    APLA_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(max_seqlen, F_BQ), num_seqs * H);
  attn_fwd_tc_kernel<<<grid, kThreads, F_SMEM, stream>>>(tq, tkv, reinterpret_cast<__nv_bfloat16*>(out), lse, cu_seqlens,
                                                         max_seqlen, H, scale);
  APLA_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace apla
